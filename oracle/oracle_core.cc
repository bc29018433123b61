// oracle_core.cc -- fields, tower, groups, pairing of the CPU restatement.
// TEST INFRASTRUCTURE ONLY (see oracle.hpp header).  Every function cites the
// reference file:line it restates (paths relative to /root/reference).
#include "oracle.hpp"
#include <stdexcept>

namespace orc {

thread_local u64 g_fq_mul_count __attribute__((tls_model("initial-exec"))) = 0;

// ---------------------------------------------------------------------------
// Repr helpers
// ---------------------------------------------------------------------------
Repr repr_from_hex(const char *hex) {  // fqrepr.go:214-217 (big-endian hex string)
    Repr r = {{0, 0, 0, 0, 0, 0}};
    size_t n = strlen(hex);
    for (size_t i = 0; i < n; i++) {
        char c = hex[n - 1 - i];
        u64 v = (c >= '0' && c <= '9') ? c - '0' : (c >= 'a' && c <= 'f') ? c - 'a' + 10 : c - 'A' + 10;
        if (i / 16 < 6) r.l[i / 16] |= v << (4 * (i % 16));
    }
    return r;
}
Repr repr_from_be48(const uint8_t b[48]) {  // fqrepr.go:181-190
    Repr r;
    for (int i = 0; i < 6; i++) {
        u64 v = 0;
        for (int j = 0; j < 8; j++) v = (v << 8) | b[(5 - i) * 8 + j];
        r.l[i] = v;
    }
    return r;
}
void repr_to_be48(const Repr &r, uint8_t b[48]) {  // fqrepr.go:193-202
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 8; j++) b[(5 - i) * 8 + j] = (uint8_t)(r.l[i] >> (8 * (7 - j)));
}

// ---------------------------------------------------------------------------
// Fq constants (fq.go:17-29, fq.go:200, fq.go:305, fq2.go:167)
// ---------------------------------------------------------------------------
const Fq FQ_ZERO = fq_raw(Repr{{0, 0, 0, 0, 0, 0}});
const Fq FQ_ONE = fq_from_repr(repr_u64(1));  // R mod Q
static Fq make_neg_one() { Fq f = FQ_ONE; f.neg(); return f; }
const Fq FQ_NEG_ONE = make_neg_one();         // fq.go:200, fq2.go:190
static Repr q_minus(u64 k, unsigned shift) {
    Repr r = sub_noborrow(Q_MOD, repr_u64(k));
    for (unsigned i = 0; i < shift; i++) r.div2();
    return r;
}
const Repr Q_MINUS_3_OVER_4 = q_minus(3, 2);  // fq2.go:167
const Repr Q_MINUS_1_OVER_2 = q_minus(1, 1);  // fq.go:305

// fq.go:96-113: MSB-first over all 384 bits, squaring only once the first set bit was seen
Fq Fq::exp(const Repr &e) const {
    Fq res = FQ_ONE;
    bool found = false;
    for (int i = 383; i >= 0; i--) {
        bool b = e.bit(i);
        if (found) res.square(); else found = b;
        if (b) res.mul(*this);
    }
    return res;
}

// fq.go:224-266 binary extended Euclid in the Montgomery domain
bool Fq::inverse(Fq &out) const {
    if (is_zero()) return false;
    const Repr one = repr_u64(1);
    Repr u = n, v = Q_MOD;
    Fq b = fq_raw(R2_MOD), c = FQ_ZERO;
    while (u.cmp(one) != 0 && v.cmp(one) != 0) {
        while (u.is_even()) {
            u.div2();
            if (b.n.is_even()) b.n.div2();
            else { b.n = add_nocarry(b.n, Q_MOD); b.n.div2(); }
        }
        while (v.is_even()) {
            v.div2();
            if (c.n.is_even()) c.n.div2();
            else { c.n = add_nocarry(c.n, Q_MOD); c.n.div2(); }
        }
        if (u.cmp(v) >= 0) { u = sub_noborrow(u, v); b.sub(c); }
        else { v = sub_noborrow(v, u); c.sub(b); }
    }
    out = (u.cmp(one) == 0) ? b : c;
    return true;
}

// fq.go:203-217
bool Fq::sqrt(Fq &out) const {
    Fq a1 = exp(Q_MINUS_3_OVER_4);
    Fq a0 = a1; a0.square(); a0.mul(*this);
    if (a0 == FQ_NEG_ONE) return false;
    a1.mul(*this);
    out = a1;
    return true;
}

// ---------------------------------------------------------------------------
// Fq2 (fq2.go)
// ---------------------------------------------------------------------------
const Fq2 FQ2_ZERO = {FQ_ZERO, FQ_ZERO};
const Fq2 FQ2_ONE = {FQ_ONE, FQ_ZERO};

// fq2.go:149-158: c1 *= {1, -1}[power % 2] (a full Fq multiplication in the reference)
void Fq2::frobenius(unsigned power) { c1.mul((power % 2) ? FQ_NEG_ONE : FQ_ONE); }

Fq2 Fq2::exp(const Repr &e) const {  // fq2.go:170-187
    Fq2 res = FQ2_ONE;
    bool found = false;
    for (int i = 383; i >= 0; i--) {
        bool b = e.bit(i);
        if (found) res.square(); else found = b;
        if (b) res.mul(*this);
    }
    return res;
}

bool Fq2::sqrt(Fq2 &out) const {  // fq2.go:198-232 (Alg. 9, eprint 2012/685)
    if (is_zero()) { out = FQ2_ZERO; return true; }
    Fq2 a1 = exp(Q_MINUS_3_OVER_4);
    Fq2 alpha = a1; alpha.square(); alpha.mul(*this);
    Fq2 a0 = alpha; a0.frobenius(1); a0.mul(alpha);
    Fq2 neg1 = {FQ_NEG_ONE, FQ_ZERO};
    if (a0 == neg1) return false;
    a1.mul(*this);
    if (alpha == neg1) {
        Fq2 i = {FQ_ZERO, FQ_ONE};
        a1.mul(i);
        out = a1;
        return true;
    }
    alpha.add(FQ2_ONE);
    alpha = alpha.exp(Q_MINUS_1_OVER_2);
    alpha.mul(a1);
    out = alpha;
    return true;
}

// ---------------------------------------------------------------------------
// Frobenius tables.  The reference hard-codes Montgomery-form literals
// (fq6.go:144-208, fq12.go:122-168); each entry is (1+u)^((q^k-1)/d).  They are
// regenerated here from that definition: c_k = conj(c_{k-1}) * c_1, because
// (q^k-1)/d = q*(q^(k-1)-1)/d + (q-1)/d and x -> x^q is conjugation on Fq2.
// ---------------------------------------------------------------------------
static Repr repr_div_small(const Repr &a, u64 d) {
    Repr q; u128 rem = 0;
    for (int i = 5; i >= 0; i--) {
        u128 cur = (rem << 64) | a.l[i];
        q.l[i] = (u64)(cur / d);
        rem = cur % d;
    }
    return q;
}
struct FrobTables {
    Fq2 fq6c1[6], fq6c2[6], fq12c1[12];
    FrobTables() {
        Fq2 xi = {FQ_ONE, FQ_ONE};  // 1+u, fq6.go:139-142
        Repr qm1 = sub_noborrow(Q_MOD, repr_u64(1));
        Fq2 g3 = xi.exp(repr_div_small(qm1, 3));
        Fq2 g6 = xi.exp(repr_div_small(qm1, 6));
        Fq2 g3sq = g3; g3sq.square();  // (1+u)^((2q-2)/3)
        auto fill = [](Fq2 *t, int n, const Fq2 &c1) {
            t[0] = FQ2_ONE;
            for (int k = 1; k < n; k++) {
                Fq2 v = t[k - 1];
                v.c1.neg();  // conjugation == x^q
                v.mul(c1);
                t[k] = v;
            }
        };
        fill(fq6c1, 6, g3);
        fill(fq6c2, 6, g3sq);
        fill(fq12c1, 12, g6);
    }
};
static const FrobTables &frob() { static const FrobTables t; return t; }
const Fq2 *frob_fq6_c1() { return frob().fq6c1; }
const Fq2 *frob_fq6_c2() { return frob().fq6c2; }
const Fq2 *frob_fq12_c1() { return frob().fq12c1; }

const Fq6 FQ6_ZERO = {FQ2_ZERO, FQ2_ZERO, FQ2_ZERO};
const Fq6 FQ6_ONE = {FQ2_ONE, FQ2_ZERO, FQ2_ZERO};
const Fq12 FQ12_ONE = {FQ6_ONE, FQ6_ZERO};

void Fq6::frobenius(unsigned power) {  // fq6.go:211-218
    c0.frobenius(power); c1.frobenius(power); c2.frobenius(power);
    c1.mul(frob().fq6c1[power % 6]);
    c2.mul(frob().fq6c2[power % 6]);
}
void Fq12::frobenius(unsigned power) {  // fq12.go:171-177
    c0.frobenius(power); c1.frobenius(power);
    c1.c0.mul(frob().fq12c1[power % 12]);
    c1.c1.mul(frob().fq12c1[power % 12]);
    c1.c2.mul(frob().fq12c1[power % 12]);
}

// ---------------------------------------------------------------------------
// Scalars
// ---------------------------------------------------------------------------
const Scalar R_MOD = {{0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL}};  // fr.go:16
bool scalar_lt(const Scalar &a, const Scalar &b) {
    for (int i = 3; i >= 0; i--) { if (a.l[i] != b.l[i]) return a.l[i] < b.l[i]; }
    return false;
}
Scalar scalar_from_be32(const uint8_t b[32]) {  // frrepr.go:178-184
    Scalar s;
    for (int i = 0; i < 4; i++) {
        u64 v = 0;
        for (int j = 0; j < 8; j++) v = (v << 8) | b[(3 - i) * 8 + j];
        s.l[i] = v;
    }
    return s;
}
void scalar_to_be32(const Scalar &s, uint8_t b[32]) {  // frrepr.go:187-194
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 8; j++) b[(3 - i) * 8 + j] = (uint8_t)(s.l[i] >> (8 * (7 - j)));
}
static Scalar scalar_sub_raw(const Scalar &a, const Scalar &b) {
    Scalar o; u64 br = 0;
    for (int i = 0; i < 4; i++) o.l[i] = sub_with_borrow(a.l[i], b.l[i], br);
    return o;
}
Scalar scalar_add_mod_r(const Scalar &a, const Scalar &b) {
    Scalar o; u64 c = 0;
    for (int i = 0; i < 4; i++) o.l[i] = add_with_carry(a.l[i], b.l[i], c);
    if (c || !scalar_lt(o, R_MOD)) o = scalar_sub_raw(o, R_MOD);
    return o;
}
Scalar scalar_mul_mod_r(const Scalar &a, const Scalar &b) {  // double-and-add, plain integers mod r
    Scalar acc = {{0, 0, 0, 0}};
    for (int i = 255; i >= 0; i--) {
        acc = scalar_add_mod_r(acc, acc);
        if (b.bit(i)) acc = scalar_add_mod_r(acc, a);
    }
    return acc;
}

// ---------------------------------------------------------------------------
// Group law, shared by G1 (over Fq) and G2 (over Fq2): g1.go:343-559 and
// g2.go:389-606 are the same formulas over different fields.
// ---------------------------------------------------------------------------
template <class F> struct Jac { F x, y, z; };

template <class F> static void jac_double(F &X, F &Y, F &Z) {  // g1.go:343-397 / g2.go:389-443
    if (Z.is_zero()) return;
    F a = X; a.square();
    F b = Y; b.square();
    F c = b; c.square();
    F d = X; d.add(b); d.square(); d.sub(a); d.sub(c); d.dbl();
    F e = a; e.dbl(); e.add(a);
    F f = e; f.square();
    F nz = Z; nz.mul(Y); nz.dbl();
    F nx = f; nx.sub(d); nx.sub(d);
    c.dbl(); c.dbl(); c.dbl();
    F ny = d; ny.sub(nx); ny.mul(e); ny.sub(c);
    X = nx; Y = ny; Z = nz;
}
template <class F> static void jac_add(F &X1, F &Y1, F &Z1, const F &X2, const F &Y2, const F &Z2) {  // g1.go:400-482
    if (Z1.is_zero()) { X1 = X2; Y1 = Y2; Z1 = Z2; return; }
    if (Z2.is_zero()) return;
    F z1z1 = Z1; z1z1.square();
    F z2z2 = Z2; z2z2.square();
    F u1 = X1; u1.mul(z2z2);
    F u2 = X2; u2.mul(z1z1);
    F s1 = Y1; s1.mul(Z2); s1.mul(z2z2);
    F s2 = Y2; s2.mul(Z1); s2.mul(z1z1);
    if (u1 == u2 && s1 == s2) { jac_double(X1, Y1, Z1); return; }
    F h = u2; h.sub(u1);
    F i = h; i.dbl(); i.square();
    F j = h; j.mul(i);
    s2.sub(s1); s2.dbl();  // r
    u1.mul(i);             // V
    F nx = s2; nx.square(); nx.sub(j); nx.sub(u1); nx.sub(u1);
    u1.sub(nx); u1.mul(s2);
    s1.mul(j); s1.dbl();
    u1.sub(s1);
    F nz = Z1; nz.add(Z2); nz.square(); nz.sub(z1z1); nz.sub(z2z2); nz.mul(h);
    X1 = nx; Y1 = u1; Z1 = nz;
}
template <class F> static void jac_add_affine(F &X1, F &Y1, F &Z1, const F &x2, const F &y2, bool inf2, const F &one) {  // g1.go:485-559
    if (Z1.is_zero()) {
        if (inf2) return;  // other.ToProjective() of zero is the canonical zero; g already is a zero
        X1 = x2; Y1 = y2; Z1 = one; return;
    }
    if (inf2) return;
    F z1z1 = Z1; z1z1.square();
    F u2 = x2; u2.mul(z1z1);
    F s2 = y2; s2.mul(Z1); s2.mul(z1z1);
    if (X1 == u2 && Y1 == s2) { jac_double(X1, Y1, Z1); return; }
    u2.sub(X1);  // H
    F hh = u2; hh.square();
    F i = hh; i.dbl(); i.dbl();
    F j = u2; j.mul(i);
    s2.sub(Y1); s2.dbl();  // r
    F v = X1; v.mul(i);
    F nx = s2; nx.square(); nx.sub(j); nx.sub(v); nx.sub(v);
    F ny = v; ny.sub(nx); ny.mul(s2);
    F i0 = Y1; i0.mul(j); i0.dbl();
    ny.sub(i0);
    F nz = Z1; nz.add(u2); nz.square(); nz.sub(z1z1); nz.sub(hh);
    X1 = nx; Y1 = ny; Z1 = nz;
}

// small adapters so Fq/Fq2 expose the same equality the reference uses
static inline bool feq(const Fq &a, const Fq &b) { return a == b; }

// ---------------------------------------------------------------------------
// G1 (g1.go)
// ---------------------------------------------------------------------------
const Fq B_COEFF = [] { Fq f = fq_from_repr(repr_u64(4)); return f; }();  // g1.go:29 (4*R mod Q)
static Fq g1_gen_x() {
    return fq_hex("17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb");
}
static Fq g1_gen_y() {
    return fq_hex("08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1");
}
const G1Affine G1_AFFINE_ZERO = {FQ_ZERO, FQ_ONE, true};           // g1.go:22
const G1Affine G1_AFFINE_ONE = {g1_gen_x(), g1_gen_y(), false};    // g1.go:25-32 (decimal there, same numbers)
const G1Proj G1_PROJ_ZERO = {FQ_ZERO, FQ_ONE, FQ_ZERO};            // g1.go:269
const G1Proj G1_PROJ_ONE = {g1_gen_x(), g1_gen_y(), FQ_ONE};       // g1.go:272

G1Proj g1_to_proj(const G1Affine &a) {  // g1.go:59-64
    if (a.infinity) return G1_PROJ_ZERO;
    return G1Proj{a.x, a.y, FQ_ONE};
}
G1Affine g1_to_affine(const G1Proj &p) {  // g1.go:322-340 (always inverts z)
    if (p.is_zero()) return G1_AFFINE_ZERO;
    Fq zi; p.z.inverse(zi);
    Fq zi2 = zi; zi2.square();
    Fq x = p.x; x.mul(zi2);
    Fq y = p.y; y.mul(zi2); y.mul(zi);
    return G1Affine{x, y, false};
}
G1Proj g1_double(const G1Proj &p) { G1Proj r = p; jac_double(r.x, r.y, r.z); return r; }
G1Proj g1_add(const G1Proj &a, const G1Proj &b) { G1Proj r = a; jac_add(r.x, r.y, r.z, b.x, b.y, b.z); return r; }
G1Proj g1_add_affine(const G1Proj &a, const G1Affine &b) {
    G1Proj r = a;
    if (a.is_zero()) return g1_to_proj(b);  // g1.go:486-488
    jac_add_affine(r.x, r.y, r.z, b.x, b.y, b.infinity, FQ_ONE);
    return r;
}
template <class BITS> static G1Proj g1_affine_mul_bits(const G1Affine &g, const BITS &b) {  // g1.go:67-90
    G1Proj res = G1_PROJ_ZERO;
    unsigned n = b.bitlen();
    for (unsigned i = 0; i < n; i++) {
        bool o = b.bit(n - i - 1);
        res = g1_double(res);
        if (o) res = g1_add_affine(res, g);
    }
    return res;
}
G1Proj g1_affine_mul_repr(const G1Affine &g, const Repr &b) { return g1_affine_mul_bits(g, b); }
G1Proj g1_affine_mul_fr(const G1Affine &g, const Scalar &b) { return g1_affine_mul_bits(g, b); }
G1Proj g1_proj_mul_fr(const G1Proj &g, const Scalar &b) {  // g1.go:575-585
    G1Proj res = G1_PROJ_ZERO;
    unsigned n = b.bitlen();
    for (unsigned i = 0; i < n; i++) {
        bool o = b.bit(n - i - 1);
        res = g1_double(res);
        if (o) res = g1_add(res, g);
    }
    return res;
}
bool g1_proj_equal(const G1Proj &a, const G1Proj &b) {  // g1.go:292-319
    if (a.is_zero()) return b.is_zero();
    if (b.is_zero()) return false;
    Fq z1 = a.z; z1.square();
    Fq z2 = b.z; z2.square();
    Fq t1 = a.x; t1.mul(z2);
    Fq t2 = b.x; t2.mul(z1);
    if (!(t1 == t2)) return false;
    z1.mul(a.z); z1.mul(b.y);
    z2.mul(b.z); z2.mul(a.y);
    return z1 == z2;
}
bool g1_is_on_curve(const G1Affine &a) {  // g1.go:93-105
    if (a.infinity) return true;
    Fq y2 = a.y; y2.square();
    Fq x3b = a.x; x3b.square(); x3b.mul(a.x); x3b.add(B_COEFF);
    return y2 == x3b;
}
bool g1_in_subgroup(const G1Affine &a) {  // g1.go:137-141 (the first MulFR result is discarded)
    return g1_affine_mul_fr(a, R_MOD).is_zero();
}
bool g1_from_x(const Fq &x, bool greatest, G1Affine &out) {  // g1.go:111-131
    Fq x3b = x; x3b.square(); x3b.mul(x); x3b.add(B_COEFF);
    Fq y;
    if (!x3b.sqrt(y)) return false;
    Fq ny = y; ny.neg();
    Fq yv = ny;
    if ((y.cmp(ny) < 0) != greatest) yv = y;
    out = G1Affine{x, yv, false};
    return true;
}
void g1_compress(const G1Affine &a, uint8_t out[48]) {  // g1.go:230-249
    memset(out, 0, 48);
    if (a.infinity) out[0] |= 1 << 6;
    else {
        repr_to_be48(a.x.to_repr(), out);
        Fq ny = a.y; ny.neg();
        if (a.y.cmp(ny) > 0) out[0] |= 1 << 5;
    }
    out[0] |= 1 << 7;
}
int g1_decompress_unchecked(const uint8_t in[48], G1Affine &out) {  // g1.go:199-227
    uint8_t c[48]; memcpy(c, in, 48);
    if ((c[0] & (1 << 7)) == 0) return 1;
    if (c[0] & (1 << 6)) {
        c[0] &= 0x3f;
        for (int i = 0; i < 48; i++) if (c[i]) return 2;
        out = G1_AFFINE_ZERO;
        return 0;
    }
    bool greatest = (c[0] & (1 << 5)) != 0;
    c[0] &= 0x1f;
    Fq x = fq_from_repr(repr_from_be48(c));
    return g1_from_x(x, greatest, out) ? 0 : 3;
}
int g1_decompress(const uint8_t in[48], G1Affine &out) {  // g1.go:185-195
    int e = g1_decompress_unchecked(in, out);
    if (e) return e;
    return g1_in_subgroup(out) ? 0 : 4;
}

// ---------------------------------------------------------------------------
// G2 (g2.go)
// ---------------------------------------------------------------------------
const Fq2 B_COEFF_FQ2 = {B_COEFF, B_COEFF};  // g2.go:32
static Fq2 g2_gen_x() {  // g2.go:26-27, 35-39
    return Fq2{fq_hex("024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8"),
               fq_hex("13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e")};
}
static Fq2 g2_gen_y() {  // g2.go:28-29, 40-43
    return Fq2{fq_hex("0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801"),
               fq_hex("0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be")};
}
const G2Affine G2_AFFINE_ZERO = {FQ2_ZERO, FQ2_ONE, true};          // g2.go:24
const G2Affine G2_AFFINE_ONE = {g2_gen_x(), g2_gen_y(), false};     // g2.go:35-43
const G2Proj G2_PROJ_ZERO = {FQ2_ZERO, FQ2_ONE, FQ2_ZERO};          // g2.go:310
const G2Proj G2_PROJ_ONE = {g2_gen_x(), g2_gen_y(), FQ2_ONE};       // g2.go:313

G2Proj g2_to_proj(const G2Affine &a) {  // g2.go:70-76
    if (a.infinity) return G2_PROJ_ZERO;
    return G2Proj{a.x, a.y, FQ2_ONE};
}
G2Affine g2_to_affine(const G2Proj &p) {  // g2.go:365-386 (shortcut when z == 1)
    if (p.is_zero()) return G2_AFFINE_ZERO;
    if (p.z == FQ2_ONE) return G2Affine{p.x, p.y, false};
    Fq2 zi = p.z; zi.inverse();
    Fq2 zi2 = zi; zi2.square();
    Fq2 x = p.x; x.mul(zi2);
    Fq2 y = p.y; y.mul(zi2); y.mul(zi);
    return G2Affine{x, y, false};
}
G2Proj g2_double(const G2Proj &p) { G2Proj r = p; jac_double(r.x, r.y, r.z); return r; }
G2Proj g2_add(const G2Proj &a, const G2Proj &b) { G2Proj r = a; jac_add(r.x, r.y, r.z, b.x, b.y, b.z); return r; }
G2Proj g2_add_affine(const G2Proj &a, const G2Affine &b) {
    if (a.is_zero()) return g2_to_proj(b);  // g2.go:533-535
    G2Proj r = a;
    jac_add_affine(r.x, r.y, r.z, b.x, b.y, b.infinity, FQ2_ONE);
    return r;
}
template <class BITS> static G2Proj g2_affine_mul_generic(const G2Affine &g, const BITS &b) {  // g2.go:79-102
    G2Proj res = G2_PROJ_ZERO;
    unsigned n = b.bitlen();
    for (unsigned i = 0; i < n; i++) {
        bool o = b.bit(n - i - 1);
        res = g2_double(res);
        if (o) res = g2_add_affine(res, g);
    }
    return res;
}
G2Proj g2_affine_mul_repr(const G2Affine &g, const Repr &b) { return g2_affine_mul_generic(g, b); }
G2Proj g2_affine_mul_fr(const G2Affine &g, const Scalar &b) { return g2_affine_mul_generic(g, b); }
struct LimbBits {
    const u64 *l; int n;
    unsigned bitlen() const {
        for (int i = n - 1; i >= 0; i--) if (l[i]) return 64 * i + 64 - __builtin_clzll(l[i]);
        return 0;
    }
    bool bit(unsigned k) const { return (l[k / 64] >> (k % 64)) & 1; }
};
G2Proj g2_affine_mul_bits(const G2Affine &g, const u64 *limbs, int nlimbs) {  // g2.go:105-115 (MulBig)
    LimbBits b{limbs, nlimbs};
    return g2_affine_mul_generic(g, b);
}
template <class BITS> static G2Proj g2_proj_mul_generic(const G2Proj &g, const BITS &b) {  // g2.go:609-632
    G2Proj res = G2_PROJ_ZERO;
    unsigned n = b.bitlen();
    for (unsigned i = 0; i < n; i++) {
        bool o = b.bit(n - i - 1);
        res = g2_double(res);
        if (o) res = g2_add(res, g);
    }
    return res;
}
G2Proj g2_proj_mul_repr(const G2Proj &g, const Repr &b) { return g2_proj_mul_generic(g, b); }
G2Proj g2_proj_mul_fr(const G2Proj &g, const Scalar &b) { return g2_proj_mul_generic(g, b); }
bool g2_proj_equal(const G2Proj &a, const G2Proj &b) {  // g2.go:333-362
    if (a.is_zero()) return b.is_zero();
    if (b.is_zero()) return false;
    Fq2 z1 = a.z; z1.square();
    Fq2 z2 = b.z; z2.square();
    Fq2 t1 = a.x; t1.mul(z2);
    Fq2 t2 = b.x; t2.mul(z1);
    if (!(t1 == t2)) return false;
    z1.mul(a.z); z1.mul(b.y);
    z2.mul(b.z); z2.mul(a.y);
    return z1 == z2;
}
bool g2_is_on_curve(const G2Affine &a) {  // g2.go:118-130
    if (a.infinity) return true;
    Fq2 y2 = a.y; y2.square();
    Fq2 x3b = a.x; x3b.square(); x3b.mul(a.x); x3b.add(B_COEFF_FQ2);
    return y2 == x3b;
}
bool g2_in_subgroup(const G2Affine &a) { return g2_affine_mul_fr(a, R_MOD).is_zero(); }  // g2.go:293-295
bool g2_from_x(const Fq2 &x, bool greatest, G2Affine &out) {  // g2.go:149-169
    Fq2 x3b = x; x3b.square(); x3b.mul(x); x3b.add(B_COEFF_FQ2);
    Fq2 y;
    if (!x3b.sqrt(y)) return false;
    Fq2 ny = y; ny.neg();
    Fq2 yv = ny;
    if ((y.cmp(ny) < 0) != greatest) yv = y;
    out = G2Affine{x, yv, false};
    return true;
}
void g2_compress(const G2Affine &a, uint8_t out[96]) {  // g2.go:268-289 (x.c1 first)
    memset(out, 0, 96);
    if (a.infinity) out[0] |= 1 << 6;
    else {
        repr_to_be48(a.x.c1.to_repr(), out);
        repr_to_be48(a.x.c0.to_repr(), out + 48);
        Fq2 ny = a.y; ny.neg();
        if (a.y.cmp(ny) > 0) out[0] |= 1 << 5;
    }
    out[0] |= 1 << 7;
}
int g2_decompress_unchecked(const uint8_t in[96], G2Affine &out) {  // g2.go:232-265
    uint8_t c[96]; memcpy(c, in, 96);
    if ((c[0] & (1 << 7)) == 0) return 1;
    if (c[0] & (1 << 6)) {
        c[0] &= 0x3f;
        for (int i = 0; i < 96; i++) if (c[i]) return 2;
        out = G2_AFFINE_ZERO;
        return 0;
    }
    bool greatest = (c[0] & (1 << 5)) != 0;
    c[0] &= 0x1f;
    Fq2 x = {fq_from_repr(repr_from_be48(c + 48)), fq_from_repr(repr_from_be48(c))};
    return g2_from_x(x, greatest, out) ? 0 : 3;
}
int g2_decompress(const uint8_t in[96], G2Affine &out) {  // g2.go:219-229
    int e = g2_decompress_unchecked(in, out);
    if (e) return e;
    return g2_in_subgroup(out) ? 0 : 4;
}

// ---------------------------------------------------------------------------
// Pairing (g2.go:634-801, pairing.go)
// ---------------------------------------------------------------------------
const u64 BLS_X = 0xd201000000010000ULL;  // g2.go:634; blsIsNegative = true (g2.go:636)

static void doubling_step(G2Proj &r, Fq2 out[3]) {  // g2.go:655-708
    Fq2 tmp0 = r.x; tmp0.square();
    Fq2 tmp1 = r.y; tmp1.square();
    Fq2 tmp2 = tmp1; tmp2.square();
    Fq2 tmp3 = tmp1; tmp3.add(r.x); tmp3.square(); tmp3.sub(tmp0); tmp3.sub(tmp2); tmp3.dbl();
    Fq2 tmp4 = tmp0; tmp4.dbl(); tmp4.add(tmp0);
    Fq2 tmp6 = r.x; tmp6.add(tmp4);
    Fq2 tmp5 = tmp4; tmp5.square();
    Fq2 zsq = r.z; zsq.square();
    r.x = tmp5; r.x.sub(tmp3); r.x.sub(tmp3);
    r.z.add(r.y); r.z.square(); r.z.sub(tmp1); r.z.sub(zsq);
    r.y = tmp3; r.y.sub(r.x); r.y.mul(tmp4);
    tmp2.dbl(); tmp2.dbl(); tmp2.dbl();
    r.y.sub(tmp2);
    tmp3 = tmp4; tmp3.mul(zsq); tmp3.dbl(); tmp3.neg();
    tmp6.square(); tmp6.sub(tmp0); tmp6.sub(tmp5);
    tmp1.dbl(); tmp1.dbl();
    tmp6.sub(tmp1);
    tmp0 = r.z; tmp0.mul(zsq); tmp0.dbl();
    out[0] = tmp0; out[1] = tmp3; out[2] = tmp6;
}
static void addition_step(G2Proj &r, const G2Affine &q, Fq2 out[3]) {  // g2.go:710-772
    Fq2 zsq = r.z; zsq.square();
    Fq2 ysq = q.y; ysq.square();
    Fq2 t0 = zsq; t0.mul(q.x);
    Fq2 t1 = q.y; t1.add(r.z); t1.square(); t1.sub(ysq); t1.sub(zsq); t1.mul(zsq);
    Fq2 t2 = t0; t2.sub(r.x);
    Fq2 t3 = t2; t3.square();
    Fq2 t4 = t3; t4.dbl(); t4.dbl();
    Fq2 t5 = t4; t5.mul(t2);
    Fq2 t6 = t1; t6.sub(r.y); t6.sub(r.y);
    Fq2 t9 = t6; t9.mul(q.x);
    Fq2 t7 = t4; t7.mul(r.x);
    r.x = t6; r.x.square(); r.x.sub(t5); r.x.sub(t7); r.x.sub(t7);
    r.z.add(t2); r.z.square(); r.z.sub(zsq); r.z.sub(t3);
    Fq2 t10 = q.y; t10.add(r.z);
    Fq2 t8 = t7; t8.sub(r.x); t8.mul(t6);
    t0 = r.y; t0.mul(t5); t0.dbl();
    r.y = t8; r.y.sub(t0);
    t10.square(); t10.sub(ysq);
    zsq = r.z; zsq.square();
    t10.sub(zsq);
    t9.dbl(); t9.sub(t10);
    t10 = r.z; t10.dbl();
    t6.neg(); t6.dbl();
    out[0] = t10; out[1] = t6; out[2] = t9;
}

G2Prepared g2_prepare(const G2Affine &q) {  // g2.go:650-801
    G2Prepared p;
    p.infinity = q.infinity;
    if (q.infinity) return p;
    G2Proj r = g2_to_proj(q);
    const u64 xr = BLS_X >> 1;
    const unsigned bl = 64 - __builtin_clzll(xr);
    bool found = false;
    Fq2 o[3];
    for (unsigned i = 0; i <= bl; i++) {
        unsigned bi = bl - i;
        bool set = bi < 64 ? ((xr >> bi) & 1) : false;
        if (!found) { found = set; continue; }
        doubling_step(r, o);
        p.coeffs.insert(p.coeffs.end(), o, o + 3);
        if (set) {
            addition_step(r, q, o);
            p.coeffs.insert(p.coeffs.end(), o, o + 3);
        }
    }
    doubling_step(r, o);
    p.coeffs.insert(p.coeffs.end(), o, o + 3);
    return p;
}

static void ell(Fq12 &f, const Fq2 *coeffs, const G1Affine &p) {  // pairing.go:28-39
    Fq2 c0 = coeffs[0], c1 = coeffs[1];
    c0.c0.mul(p.y); c0.c1.mul(p.y);
    c1.c0.mul(p.x); c1.c1.mul(p.x);
    f.mul_by_014(coeffs[2], c1, c0);
}

Fq12 miller_loop(const std::vector<MillerItem> &items) {  // pairing.go:16-75
    struct Pair { G1Affine p; const Fq2 *q; size_t idx; };
    std::vector<Pair> pairs;
    for (const MillerItem &it : items) {
        // The reference leaves a zero-valued pair for infinity inputs and then panics on it
        // (SURVEY.md Q1).  The oracle skips such pairs (factor 1), which is what the C ABI documents.
        if (!it.p.infinity && !it.q->infinity) pairs.push_back(Pair{it.p, it.q->coeffs.data(), 0});
    }
    Fq12 f = FQ12_ONE;
    const u64 xr = BLS_X >> 1;
    const unsigned bl = 64 - __builtin_clzll(xr);
    bool found = false;
    for (unsigned q = 0; q <= bl; q++) {
        unsigned bi = bl - q;
        bool set = bi < 64 ? ((xr >> bi) & 1) : false;
        if (!found) { found = set; continue; }
        for (Pair &pr : pairs) { ell(f, pr.q + 3 * pr.idx, pr.p); pr.idx++; }
        if (set) for (Pair &pr : pairs) { ell(f, pr.q + 3 * pr.idx, pr.p); pr.idx++; }
        f.square();
    }
    for (Pair &pr : pairs) { ell(f, pr.q + 3 * pr.idx, pr.p); pr.idx++; }
    f.conjugate();  // blsIsNegative
    return f;
}

static Fq12 exp_by_x(const Fq12 &f, u64 x) {  // pairing.go:92-98
    Fq12 r = f.exp(repr_u64(x));
    r.conjugate();
    return r;
}

bool final_exponentiation(const Fq12 &r_in, Fq12 &out) {  // pairing.go:79-129
    Fq12 f1 = r_in; f1.conjugate();
    Fq12 f2 = r_in;
    if (!f2.inverse()) return false;
    Fq12 r = f1; r.mul(f2);
    f2 = r;
    r.frobenius(2);
    r.mul(f2);
    u64 x = BLS_X;
    Fq12 y0 = r; y0.square();
    Fq12 y1 = exp_by_x(y0, x);
    x >>= 1;
    Fq12 y2 = exp_by_x(y1, x);
    x <<= 1;
    Fq12 y3 = r; y3.conjugate();
    y1.mul(y3);
    y1.conjugate();
    y1.mul(y2);
    y2 = exp_by_x(y1, x);
    y3 = exp_by_x(y2, x);
    y1.conjugate();
    y3.mul(y1);
    y1.conjugate();
    y1.frobenius(3);
    y2.frobenius(2);
    y1.mul(y2);
    y2 = exp_by_x(y3, x);
    y2.mul(y0);
    y2.mul(r);
    y1.mul(y2);
    y3.frobenius(1);
    y1.mul(y3);
    out = y1;
    return true;
}

Fq12 pairing(const G1Proj &p, const G2Proj &q) {  // pairing.go:132-136
    G2Prepared prep = g2_prepare(g2_to_affine(q));
    std::vector<MillerItem> items{MillerItem{g1_to_affine(p), &prep}};
    Fq12 out = FQ12_ONE;
    final_exponentiation(miller_loop(items), out);
    return out;
}

bool compare_two_pairings(const G1Proj &p1, const G2Proj &q1, const G1Proj &p2, const G2Proj &q2) {  // pairing.go:140-147
    G1Proj np2 = p2; np2.neg();
    G2Prepared a = g2_prepare(g2_to_affine(q1));
    G2Prepared b = g2_prepare(g2_to_affine(q2));
    std::vector<MillerItem> items{MillerItem{g1_to_affine(p1), &a}, MillerItem{g1_to_affine(np2), &b}};
    Fq12 out;
    if (!final_exponentiation(miller_loop(items), out)) return false;
    return out == FQ12_ONE;
}

// ---------------------------------------------------------------------------
// Test RNG: Go crypto/rand.Int over the xorshift reader
// ---------------------------------------------------------------------------
Scalar rand_fr(XorShift &r) {  // fr.go:337-344: k = 32 bytes, top-byte mask 0x7f, reject >= r
    for (;;) {
        uint8_t b[32];
        for (int i = 0; i < 32; i++) b[i] = r.next_byte();
        b[0] &= 0x7f;
        Scalar s = scalar_from_be32(b);
        if (scalar_lt(s, R_MOD)) return s;
    }
}
Fq rand_fq(XorShift &r) {  // fq.go:341-349: k = 48 bytes, top-byte mask 0x1f, reject >= Q
    for (;;) {
        uint8_t b[48];
        for (int i = 0; i < 48; i++) b[i] = r.next_byte();
        b[0] &= 0x1f;
        Repr v = repr_from_be48(b);
        if (v.cmp(Q_MOD) < 0) return fq_from_repr(v);
    }
}

}  // namespace orc
