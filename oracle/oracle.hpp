// oracle.hpp -- CPU restatement of the phoreproject/bls pure-Go path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product (bls_b200/, include/) may
// include, link or call this file; only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs use it, as the checker and the
// timed CPU baseline.  The Go toolchain is absent from this image, so the
// reference itself cannot run here: this restatement follows the reference's
// algorithms function by function (each cites file:line under /root/reference)
// and is pinned by the reference's own golden vectors (oracle/selftest.cc,
// tests/golden/).
//
// Representation follows the reference: Fq = 6 x u64 limbs, least significant
// first (fqrepr.go:13-14), Montgomery form with R = 2^384, always canonical in
// [0,Q) (fq.go:41-45).
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

namespace orc {

typedef uint64_t u64;
typedef unsigned __int128 u128;

// per-thread count of Fq multiplications + squarings (SURVEY.md section 8d unit of work)
extern thread_local u64 g_fq_mul_count __attribute__((tls_model("initial-exec")));

// ---------------------------------------------------------------------------
// L0 limb primitives -- stub_fallback.go:11-155
// ---------------------------------------------------------------------------
// stub_fallback.go:149-155
static inline u64 mac_with_carry(u64 a, u64 b, u64 c, u64 &carry) {
    u128 t = (u128)b * c + a + carry;
    carry = (u64)(t >> 64);
    return (u64)t;
}
// stub_fallback.go:136-140
static inline u64 add_with_carry(u64 a, u64 b, u64 &carry) {
    u128 t = (u128)a + b + carry;
    carry = (u64)(t >> 64);
    return (u64)t;
}
// stub_fallback.go:144-147
static inline u64 sub_with_borrow(u64 a, u64 b, u64 &borrow) {
    u128 t = (u128)a - b - borrow;
    borrow = (u64)(t >> 64) & 1;
    return (u64)t;
}

struct Repr {  // FQRepr, fqrepr.go:13-14
    u64 l[6];
    bool operator==(const Repr &o) const { return memcmp(l, o.l, sizeof l) == 0; }
    bool is_zero() const { return (l[0] | l[1] | l[2] | l[3] | l[4] | l[5]) == 0; }
    bool is_even() const { return (l[0] & 1) == 0; }
    int cmp(const Repr &g) const {  // fqrepr.go:143-156
        for (int i = 5; i >= 0; i--) {
            if (l[i] == g.l[i]) continue;
            return l[i] > g.l[i] ? 1 : -1;
        }
        return 0;
    }
    void div2() {  // fqrepr.go:70-78
        u64 t = 0;
        for (int i = 5; i >= 0; i--) { u64 t2 = l[i] << 63; l[i] >>= 1; l[i] |= t; t = t2; }
    }
    void mul2() {  // fqrepr.go:81-89
        u64 last = 0;
        for (int i = 0; i < 6; i++) { u64 tmp = l[i] >> 63; l[i] <<= 1; l[i] |= last; last = tmp; }
    }
    unsigned bitlen() const {  // fqrepr.go:169-179
        for (int i = 5; i >= 0; i--) if (l[i]) return 64 * i + 64 - __builtin_clzll(l[i]);
        return 0;
    }
    bool bit(unsigned n) const { return (l[n / 64] >> (n % 64)) & 1; }  // fqrepr.go:205-207
};

// stub_fallback.go:119-126
static inline Repr add_nocarry(const Repr &a, const Repr &b) {
    Repr o; u64 c = 0;
    for (int i = 0; i < 6; i++) o.l[i] = add_with_carry(a.l[i], b.l[i], c);
    return o;
}
// stub_fallback.go:128-134
static inline Repr sub_noborrow(const Repr &a, const Repr &b) {
    Repr o; u64 br = 0;
    for (int i = 0; i < 6; i++) o.l[i] = sub_with_borrow(a.l[i], b.l[i], br);
    return o;
}

// Q, fq.go:26 (limbs: asm/asm.go:40)
static const Repr Q_MOD = {{0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                            0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL}};
// R^2 mod Q, fq.go:29
static const Repr R2_MOD = {{0xf4df1f341c341746ULL, 0x0a76e6a609d104f1ULL, 0x8de5476c4c95b6d5ULL,
                             0x67eb88a9939d83c0ULL, 0x9a793e85b519952dULL, 0x11988fe592cae3aaULL}};
static const u64 MONT_INV = 0x89f3fffcfffcfffdULL;  // stub_fallback.go:59

// stub_fallback.go:11-57 (6x6 schoolbook, row by row)
static inline void multiply_fq_repr(const u64 a[6], const u64 b[6], u64 hi[6], u64 lo[6]) {
    u64 r[12] = {0};
    for (int i = 0; i < 6; i++) {
        u64 carry = 0;
        for (int j = 0; j < 6; j++) r[i + j] = mac_with_carry(r[i + j], a[i], b[j], carry);
        r[i + 6] = carry;
    }
    for (int i = 0; i < 6; i++) { lo[i] = r[i]; hi[i] = r[i + 6]; }
}

// stub_fallback.go:61-116 (6 rounds, carry + carry2 chains, final carry dropped)
static inline Repr mont_reduce(const u64 hi_in[6], const u64 lo_in[6]) {
    u64 r[12];
    for (int i = 0; i < 6; i++) { r[i] = lo_in[i]; r[i + 6] = hi_in[i]; }
    u64 carry2 = 0;
    for (int i = 0; i < 6; i++) {
        u64 k = r[i] * MONT_INV;
        u64 carry = 0;
        (void)mac_with_carry(r[i], k, Q_MOD.l[0], carry);
        for (int j = 1; j < 6; j++) r[i + j] = mac_with_carry(r[i + j], k, Q_MOD.l[j], carry);
        // hi[i], carry = AddWithCarry(hi[i], carry2, carry)
        u128 t = (u128)r[i + 6] + carry2 + carry;
        r[i + 6] = (u64)t;
        carry2 = (u64)(t >> 64);
    }
    Repr o;
    for (int i = 0; i < 6; i++) o.l[i] = r[i + 6];
    return o;
}

// ---------------------------------------------------------------------------
// Fq -- fq.go
// ---------------------------------------------------------------------------
struct Fq {
    Repr n;
    bool operator==(const Fq &o) const { return n == o.n; }  // fq.go:116-118
    bool is_zero() const { return n.is_zero(); }               // fq.go:146-148
    bool is_valid() const {                                    // fq.go:37-39
        return (n.l[5] & 0xf000000000000000ULL) == 0 || n.cmp(Q_MOD) < 0;
    }
    void reduce() { if (!is_valid()) n = sub_noborrow(n, Q_MOD); }  // fq.go:41-45
    void add(const Fq &o) { n = add_nocarry(n, o.n); reduce(); }    // fq.go:65-68
    void mont_red(const u64 hi[6], const u64 lo[6]) { n = mont_reduce(hi, lo); reduce(); }  // fq.go:70-73
    void mul(const Fq &o) {                                        // fq.go:76-79
        g_fq_mul_count++;
        u64 hi[6], lo[6];
        multiply_fq_repr(n.l, o.n.l, hi, lo);
        mont_red(hi, lo);
    }
    void sub(const Fq &o) {                                        // fq.go:82-87
        if (o.n.cmp(n) > 0) n = add_nocarry(n, Q_MOD);
        n = sub_noborrow(n, o.n);
    }
    void neg() { if (!is_zero()) n = sub_noborrow(Q_MOD, n); }     // fq.go:121-127
    void dbl() { n.mul2(); reduce(); }                             // fq.go:140-143
    void square() {                                                // fq.go:151-198
        g_fq_mul_count++;
        const u64 *f = n.l;
        u64 r[12] = {0};
        // off-diagonal products, one row per i (fq.go:152-172)
        for (int i = 0; i < 5; i++) {
            u64 carry = 0;
            for (int j = i + 1; j < 6; j++) r[i + j] = mac_with_carry(r[i + j], f[i], f[j], carry);
            r[i + 6] = carry;
        }
        // shift left by one bit (fq.go:173-183)
        r[11] = r[10] >> 63;
        for (int i = 10; i >= 2; i--) r[i] = (r[i] << 1) | (r[i - 1] >> 63);
        r[1] = r[1] << 1;
        // diagonal (fq.go:185-197)
        u64 carry = 0;
        for (int i = 0; i < 6; i++) {
            r[2 * i] = mac_with_carry(r[2 * i], f[i], f[i], carry);
            r[2 * i + 1] = add_with_carry(r[2 * i + 1], 0, carry);
        }
        mont_red(r + 6, r);
    }
    Repr to_repr() const {                                          // fq.go:334-338
        u64 z[6] = {0, 0, 0, 0, 0, 0};
        Fq o; o.mont_red(z, n.l);
        return o.n;
    }
    int cmp(const Fq &o) const { return to_repr().cmp(o.to_repr()); }  // fq.go:134-137
    bool parity() const { Fq ng = *this; ng.neg(); return cmp(ng) > 0; }  // fq.go:269-273
    // fq.go:96-113 (MSB-first over all 384 bits; bititerator.go)
    Fq exp(const Repr &e) const;
    bool inverse(Fq &out) const;     // fq.go:224-266
    bool sqrt(Fq &out) const;        // fq.go:203-217
};

static inline Fq fq_raw(const Repr &r) { Fq f; f.n = r; return f; }     // fq.go:60-62
static inline Fq fq_from_repr(const Repr &r) {                            // fq.go:49-56
    Fq f; f.n = r;
    if (f.is_valid()) { f.mul(fq_raw(R2_MOD)); return f; }
    return fq_raw(Repr{{0, 0, 0, 0, 0, 0}});
}
static inline Repr repr_u64(u64 v) { return Repr{{v, 0, 0, 0, 0, 0}}; }
Repr repr_from_hex(const char *hex);                 // fqrepr.go:214-217
Repr repr_from_be48(const uint8_t b[48]);            // fqrepr.go:181-190
void repr_to_be48(const Repr &r, uint8_t b[48]);     // fqrepr.go:193-202
static inline Fq fq_hex(const char *h) { return fq_from_repr(repr_from_hex(h)); }

extern const Fq FQ_ZERO, FQ_ONE, FQ_NEG_ONE;       // fq.go:17-23, fq.go:200
extern const Repr Q_MINUS_3_OVER_4, Q_MINUS_1_OVER_2;  // fq2.go:167, fq.go:305

// ---------------------------------------------------------------------------
// Fq2 -- fq2.go
// ---------------------------------------------------------------------------
struct Fq2 {
    Fq c0, c1;
    bool is_zero() const { return c0.is_zero() && c1.is_zero(); }  // fq2.go:70-72
    int cmp(const Fq2 &o) const {                                  // fq2.go:31-37
        int c = c1.cmp(o.c1);
        return c != 0 ? c : c0.cmp(o.c0);
    }
    bool operator==(const Fq2 &o) const { return cmp(o) == 0; }    // fq2.go:193-195
    void mul_by_nonresidue() {                                     // fq2.go:41-45
        Fq old = c0; c0.sub(c1); c1.add(old);
    }
    Fq norm() const { Fq t0 = c0, t1 = c1; t0.square(); t1.square(); t1.add(t0); return t1; }  // fq2.go:48-55
    void square() {                                                // fq2.go:75-89
        Fq ab = c0; ab.mul(c1);
        Fq c0c1 = c0; c0c1.add(c1);
        Fq t = c1; t.neg(); t.add(c0);
        t.mul(c0c1);
        t.sub(ab); t.add(ab);
        ab.add(ab);
        c0 = t; c1 = ab;
    }
    void dbl() { c0.dbl(); c1.dbl(); }                             // fq2.go:92-95
    void neg() { c0.neg(); c1.neg(); }                             // fq2.go:98-101
    void add(const Fq2 &o) { c0.add(o.c0); c1.add(o.c1); }         // fq2.go:104-107
    void sub(const Fq2 &o) { c0.sub(o.c0); c1.sub(o.c1); }         // fq2.go:110-113
    void mul(const Fq2 &o) {                                       // fq2.go:116-130
        Fq aa = c0; aa.mul(o.c0);
        Fq bb = c1; bb.mul(o.c1);
        Fq s = o.c0; s.add(o.c1);
        c1.add(c0); c1.mul(s); c1.sub(aa); c1.sub(bb);
        c0 = aa; c0.sub(bb);
    }
    bool inverse() {                                               // fq2.go:133-147
        Fq t1 = c1; t1.square();
        Fq t0 = c0; t0.square();
        t0.add(t1);
        Fq t;
        if (!t0.inverse(t)) return false;
        c0.mul(t); c1.mul(t); c1.neg();
        return true;
    }
    void frobenius(unsigned power);                                // fq2.go:149-158
    Fq2 exp(const Repr &e) const;                                  // fq2.go:170-187
    bool sqrt(Fq2 &out) const;                                     // fq2.go:198-232
    bool parity() const { Fq2 ng = *this; ng.neg(); return cmp(ng) > 0; }  // fq2.go (Parity)
};
extern const Fq2 FQ2_ZERO, FQ2_ONE;

// ---------------------------------------------------------------------------
// Fq6 -- fq6.go
// ---------------------------------------------------------------------------
struct Fq6 {
    Fq2 c0, c1, c2;
    bool operator==(const Fq6 &o) const { return c0 == o.c0 && c1 == o.c1 && c2 == o.c2; }  // fq6.go:99-101
    void mul_by_nonresidue() {                                     // fq6.go:34-37
        Fq2 o0 = c0, o1 = c1, o2 = c2;
        c0 = o2; c1 = o0; c2 = o1;
        c0.mul_by_nonresidue();
    }
    void mul_by_1(const Fq2 &d1) {                                 // fq6.go:40-57
        Fq2 b = c1; b.mul(d1);
        Fq2 tmp = c1; tmp.add(c2);
        Fq2 t1 = d1; t1.mul(tmp); t1.sub(b); t1.mul_by_nonresidue();
        tmp = c0; tmp.add(c1);
        Fq2 t2 = d1; t2.mul(tmp); t2.sub(b);
        c0 = t1; c1 = t2; c2 = b;
    }
    void mul_by_01(const Fq2 &d0, const Fq2 &d1) {                 // fq6.go:60-90
        Fq2 a = c0; a.mul(d0);
        Fq2 b = c1; b.mul(d1);
        Fq2 tmp = c1; tmp.add(c2);
        Fq2 t1 = d1; t1.mul(tmp); t1.sub(b); t1.mul_by_nonresidue(); t1.add(a);
        tmp = c0; tmp.add(c2);
        Fq2 t3 = d0; t3.mul(tmp); t3.sub(a); t3.add(b);
        tmp = c0; tmp.add(c1);
        Fq2 t2 = d0; t2.add(d1); t2.mul(tmp); t2.sub(a); t2.sub(b);
        c0 = t1; c1 = t2; c2 = t3;
    }
    void dbl() { c0.dbl(); c1.dbl(); c2.dbl(); }                   // fq6.go:109-113
    void neg() { c0.neg(); c1.neg(); c2.neg(); }                   // fq6.go:116-120
    void add(const Fq6 &o) { c0.add(o.c0); c1.add(o.c1); c2.add(o.c2); }  // fq6.go:123-127
    void sub(const Fq6 &o) { c0.sub(o.c0); c1.sub(o.c1); c2.sub(o.c2); }  // fq6.go:130-134
    void frobenius(unsigned power);                                // fq6.go:211-218
    void square() {                                                // fq6.go:221-252
        Fq2 s0 = c0; s0.square();
        Fq2 ab = c0; ab.mul(c1);
        Fq2 s1 = ab; s1.dbl();
        Fq2 s2 = c0; s2.sub(c1); s2.add(c2); s2.square();
        Fq2 bc = c1; bc.mul(c2);
        Fq2 s3 = bc; s3.dbl();
        Fq2 s4 = c2; s4.square();
        c0 = s3; c0.mul_by_nonresidue(); c0.add(s0);
        c1 = s4; c1.mul_by_nonresidue(); c1.add(s1);
        c2 = s1; c2.add(s2); c2.add(s3); c2.sub(s0); c2.sub(s4);
    }
    void mul(const Fq6 &o) {                                       // fq6.go:255-292
        Fq2 aa = c0; aa.mul(o.c0);
        Fq2 bb = c1; bb.mul(o.c1);
        Fq2 cc = c2; cc.mul(o.c2);
        Fq2 tmp = c1; tmp.add(c2);
        Fq2 t1 = o.c1; t1.add(o.c2); t1.mul(tmp); t1.sub(bb); t1.sub(cc); t1.mul_by_nonresidue(); t1.add(aa);
        tmp = c0; tmp.add(c2);
        Fq2 n2 = o.c0; n2.add(o.c2); n2.mul(tmp); n2.sub(aa); n2.add(bb); n2.sub(cc);
        tmp = c0; tmp.add(c1);
        Fq2 n1 = o.c0; n1.add(o.c1); n1.mul(tmp); n1.sub(aa); n1.sub(bb);
        cc.mul_by_nonresidue();
        n1.add(cc);
        c0 = t1; c1 = n1; c2 = n2;
    }
    bool inverse() {                                               // fq6.go:295-336
        Fq2 k0 = c2; k0.mul_by_nonresidue(); k0.mul(c1); k0.neg();
        Fq2 c0s = c0; c0s.square(); k0.add(c0s);
        Fq2 k1 = c2; k1.square(); k1.mul_by_nonresidue();
        Fq2 c0c1 = c0; c0c1.mul(c1);
        Fq2 c0c2 = c0; c0c2.mul(c2);
        Fq2 c1c2 = c1; c1c2.mul(c2);   // computed by the reference, unused
        k1.sub(c0c1);
        Fq2 k2 = c1; k2.square(); k2.sub(c0c2);
        Fq2 tmp1 = c2; tmp1.mul(k1);
        Fq2 tmp2 = c1; tmp2.mul(k2);
        tmp1.add(tmp2); tmp1.mul_by_nonresidue();
        tmp2 = c0; tmp2.mul(k0);
        tmp1.add(tmp2);
        if (!tmp1.inverse()) return false;
        c0 = tmp1; c0.mul(k0);
        c1 = tmp1; c1.mul(k1);
        c2 = tmp1; c2.mul(k2);
        return true;
    }
};
extern const Fq6 FQ6_ZERO, FQ6_ONE;

// ---------------------------------------------------------------------------
// Fq12 -- fq12.go
// ---------------------------------------------------------------------------
struct Fq12 {
    Fq6 c0, c1;
    bool operator==(const Fq12 &o) const { return c0 == o.c0 && c1 == o.c1; }  // fq12.go:56-58
    bool is_zero() const { return c0 == FQ6_ZERO && c1 == FQ6_ZERO; }          // fq12.go:79-81
    void conjugate() { c1.neg(); }                                             // fq12.go:27-29
    void mul_by_014(const Fq2 &d0, const Fq2 &d1_in, const Fq2 &d4) {          // fq12.go:32-47
        Fq2 d1 = d1_in;
        Fq6 aa = c0; aa.mul_by_01(d0, d1);
        Fq6 bb = c1; bb.mul_by_1(d4);
        d1.add(d4);
        c1.add(c0);
        c1.mul_by_01(d0, d1);
        c1.sub(aa); c1.sub(bb);
        c0 = bb; c0.mul_by_nonresidue(); c0.add(aa);
    }
    void frobenius(unsigned power);                                            // fq12.go:171-177
    void square() {                                                            // fq12.go:180-195
        Fq6 ab = c0; ab.mul(c1);
        Fq6 c0c1 = c0; c0c1.add(c1);
        Fq6 t = c1; t.mul_by_nonresidue(); t.add(c0);
        t.mul(c0c1);
        t.sub(ab);
        c1 = ab; c1.add(ab);
        ab.mul_by_nonresidue();
        t.sub(ab);
        c0 = t;
    }
    void mul(const Fq12 &o) {                                                  // fq12.go:198-213
        Fq6 aa = c0; aa.mul(o.c0);
        Fq6 bb = c1; bb.mul(o.c1);
        Fq6 s = o.c0; s.add(o.c1);
        c1.add(c0); c1.mul(s); c1.sub(aa); c1.sub(bb);
        c0 = bb; c0.mul_by_nonresidue(); c0.add(aa);
    }
    bool inverse() {                                                           // fq12.go:216-237
        Fq6 c0s = c0; c0s.square();
        Fq6 c1s = c1; c1s.square(); c1s.mul_by_nonresidue();
        c0s.sub(c1s);
        if (!c0s.inverse()) return false;
        Fq6 t0 = c0s; t0.mul(c0);
        Fq6 t1 = c0s; t1.mul(c1); t1.neg();
        c0 = t0; c1 = t1;
        return true;
    }
    Fq12 exp(const Repr &e) const {                                            // fq12.go:108-120 (LSB-first)
        Repr nc = e;
        Fq12 res; res.c0 = FQ6_ONE; res.c1 = FQ6_ZERO;
        Fq12 fi = *this;
        while (!nc.is_zero()) {
            if (!nc.is_even()) res.mul(fi);
            Fq12 fc = fi; fi.mul(fc);     // fi.MulAssign(fi): alias-safe in the reference, equals fi^2
            nc.div2();
        }
        return res;
    }
};
extern const Fq12 FQ12_ONE;

// ---------------------------------------------------------------------------
// Scalars (FRRepr as canonical 256-bit integers) -- frrepr.go, fr.go
// ---------------------------------------------------------------------------
struct Scalar {
    u64 l[4];
    unsigned bitlen() const {  // frrepr.go:164-175
        for (int i = 3; i >= 0; i--) if (l[i]) return 64 * i + 64 - __builtin_clzll(l[i]);
        return 0;
    }
    bool bit(unsigned n) const { return (l[n / 64] >> (n % 64)) & 1; }  // frrepr.go:197-199
    bool operator==(const Scalar &o) const { return memcmp(l, o.l, sizeof l) == 0; }
};
extern const Scalar R_MOD;  // fr.go:16
bool scalar_lt(const Scalar &a, const Scalar &b);
Scalar scalar_from_be32(const uint8_t b[32]);   // frrepr.go:178-184
void scalar_to_be32(const Scalar &s, uint8_t b[32]);  // frrepr.go:187-194
Scalar scalar_add_mod_r(const Scalar &a, const Scalar &b);
Scalar scalar_mul_mod_r(const Scalar &a, const Scalar &b);

// ---------------------------------------------------------------------------
// G1 -- g1.go
// ---------------------------------------------------------------------------
struct G1Affine { Fq x, y; bool infinity; };                 // g1.go:10-14
struct G1Proj {                                              // g1.go:252-256
    Fq x, y, z;
    bool is_zero() const { return z.is_zero(); }             // g1.go:287-289
    void neg() { y.neg(); }                                  // g1.go:259-261
};
extern const G1Affine G1_AFFINE_ZERO, G1_AFFINE_ONE;         // g1.go:22, :32
extern const G1Proj G1_PROJ_ZERO, G1_PROJ_ONE;               // g1.go:269, :272
extern const Fq B_COEFF;                                     // g1.go:29

G1Proj g1_to_proj(const G1Affine &a);                        // g1.go:59-64
G1Affine g1_to_affine(const G1Proj &p);                      // g1.go:322-340
G1Proj g1_double(const G1Proj &p);                           // g1.go:343-397
G1Proj g1_add(const G1Proj &a, const G1Proj &b);             // g1.go:400-482
G1Proj g1_add_affine(const G1Proj &a, const G1Affine &b);    // g1.go:485-559
G1Proj g1_affine_mul_repr(const G1Affine &g, const Repr &b); // g1.go:67-77
G1Proj g1_affine_mul_fr(const G1Affine &g, const Scalar &b); // g1.go:80-90
G1Proj g1_proj_mul_fr(const G1Proj &g, const Scalar &b);     // g1.go:575-585
bool g1_proj_equal(const G1Proj &a, const G1Proj &b);        // g1.go:292-319
bool g1_is_on_curve(const G1Affine &a);                      // g1.go:93-105
bool g1_in_subgroup(const G1Affine &a);                      // g1.go:137-141
bool g1_from_x(const Fq &x, bool greatest, G1Affine &out);   // g1.go:111-131
void g1_compress(const G1Affine &a, uint8_t out[48]);        // g1.go:230-249
// returns 0 ok, else error code (1 mode, 2 infinity junk, 3 not on curve, 4 subgroup)
int g1_decompress_unchecked(const uint8_t in[48], G1Affine &out);  // g1.go:199-227
int g1_decompress(const uint8_t in[48], G1Affine &out);            // g1.go:185-195

// ---------------------------------------------------------------------------
// G2 -- g2.go
// ---------------------------------------------------------------------------
struct G2Affine { Fq2 x, y; bool infinity; };               // g2.go:12-16
struct G2Proj {                                             // g2.go:298-302
    Fq2 x, y, z;
    bool is_zero() const { return z.is_zero(); }            // g2.go:328-330
};
extern const G2Affine G2_AFFINE_ZERO, G2_AFFINE_ONE;        // g2.go:24, :35-43
extern const G2Proj G2_PROJ_ZERO, G2_PROJ_ONE;              // g2.go:310, :313
extern const Fq2 B_COEFF_FQ2;                               // g2.go:32

G2Proj g2_to_proj(const G2Affine &a);                       // g2.go:70-76
G2Affine g2_to_affine(const G2Proj &p);                     // g2.go:365-386
G2Proj g2_double(const G2Proj &p);                          // g2.go:389-443
G2Proj g2_add(const G2Proj &a, const G2Proj &b);            // g2.go:446-529
G2Proj g2_add_affine(const G2Proj &a, const G2Affine &b);   // g2.go:532-606
G2Proj g2_affine_mul_repr(const G2Affine &g, const Repr &b);    // g2.go:79-89
G2Proj g2_affine_mul_fr(const G2Affine &g, const Scalar &b);    // g2.go:92-102
G2Proj g2_affine_mul_bits(const G2Affine &g, const u64 *limbs, int nlimbs);  // g2.go:105-115 (MulBig)
G2Proj g2_proj_mul_repr(const G2Proj &g, const Repr &b);        // g2.go:609-619
G2Proj g2_proj_mul_fr(const G2Proj &g, const Scalar &b);        // g2.go:622-632
G2Proj g2_scale_by_cofactor(const G2Affine &g);                 // g2.go:132-138
bool g2_proj_equal(const G2Proj &a, const G2Proj &b);           // g2.go:333-362
bool g2_is_on_curve(const G2Affine &a);                         // g2.go:118-130
bool g2_in_subgroup(const G2Affine &a);                         // g2.go:293-295
bool g2_from_x(const Fq2 &x, bool greatest, G2Affine &out);     // g2.go:149-169
void g2_compress(const G2Affine &a, uint8_t out[96]);           // g2.go:268-289
int g2_decompress_unchecked(const uint8_t in[96], G2Affine &out);  // g2.go:232-265
int g2_decompress(const uint8_t in[96], G2Affine &out);            // g2.go:219-229

// ---------------------------------------------------------------------------
// Pairing -- g2.go:634-801, pairing.go
// ---------------------------------------------------------------------------
struct G2Prepared {                                            // g2.go:639-642
    std::vector<Fq2> coeffs;  // 3 per step
    bool infinity;
    size_t nsteps() const { return coeffs.size() / 3; }
};
extern const u64 BLS_X;                                        // g2.go:634
G2Prepared g2_prepare(const G2Affine &q);                      // g2.go:650-801
struct MillerItem { G1Affine p; const G2Prepared *q; };        // pairing.go:4-7
Fq12 miller_loop(const std::vector<MillerItem> &items);        // pairing.go:16-75
bool final_exponentiation(const Fq12 &r, Fq12 &out);           // pairing.go:79-129 (false == nil)
Fq12 pairing(const G1Proj &p, const G2Proj &q);                // pairing.go:132-136
bool compare_two_pairings(const G1Proj &p1, const G2Proj &q1, const G1Proj &p2, const G2Proj &q2);  // pairing.go:140-147

// ---------------------------------------------------------------------------
// Hashing -- hash.go, g1.go:614-714, g2.go:883-1085
// ---------------------------------------------------------------------------
void sha256(const uint8_t *msg, size_t len, uint8_t out[32]);
Scalar hash_secret_key(const uint8_t b[32]);                        // hash.go:9-39
Fq hp(const uint8_t *msg, size_t len, uint8_t ctr);                 // hash.go:41-72
Fq2 hp2(const uint8_t *msg, size_t len, uint8_t ctr);               // hash.go:74-113
G1Affine hash_g1(const uint8_t *msg, size_t len);                   // hash.go:326-331
G2Affine hash_g2(const uint8_t *msg, size_t len);                   // hash.go:405-411
G2Proj hash_g2_with_domain(const uint8_t msg[32], const uint8_t domain[8]);  // g2.go:1041-1085

// ---------------------------------------------------------------------------
// Test RNG: xorshift byte reader (g1_test.go:106-124) + Go crypto/rand.Int
// ---------------------------------------------------------------------------
struct XorShift {
    u64 state;
    explicit XorShift(u64 s) : state(s) {}
    uint8_t next_byte() {
        u64 x = state;
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        state = x;
        return (uint8_t)x;
    }
};
Scalar rand_fr(XorShift &r);   // fr.go:337-344 (returns the canonical integer, i.e. FR.ToRepr())
Fq rand_fq(XorShift &r);       // fq.go:341-349

// ---------------------------------------------------------------------------
// g1pubs / g2pubs glue -- g1pubs/bls.go, g2pubs/bls.go
// ---------------------------------------------------------------------------
namespace g1pubs {
G2Proj sign(const uint8_t *msg, size_t len, const Scalar &sk);                           // :132-135
G2Proj sign_with_domain(const uint8_t msg[32], const Scalar &sk, const uint8_t dom[8]);  // :138-141
G1Proj priv_to_pub(const Scalar &sk);                                                    // :144-146
bool verify(const uint8_t *msg, size_t len, const G1Proj &pub, const G2Proj &sig);       // :165-168
bool verify_with_domain(const uint8_t msg[32], const G1Proj &pub, const G2Proj &sig, const uint8_t dom[8]);  // :171-174
G2Proj aggregate_signatures(const std::vector<G2Proj> &sigs);                            // :177-189
G1Proj aggregate_public_keys(const std::vector<G1Proj> &pubs);                           // :192-204
bool verify_aggregate(const G2Proj &sig, const std::vector<G1Proj> &pubs,
                      const std::vector<std::string> &msgs);                             // :252-282
bool verify_aggregate_common(const G2Proj &sig, const std::vector<G1Proj> &pubs,
                             const uint8_t *msg, size_t len);                            // :287-290
bool verify_aggregate_common_with_domain(const G2Proj &sig, const std::vector<G1Proj> &pubs,
                                         const uint8_t msg[32], const uint8_t dom[8]);   // :294-297
bool verify_aggregate_with_domain(const G2Proj &sig, const std::vector<G1Proj> &pubs,
                                  const std::vector<std::string> &msgs32, const uint8_t dom[8]);  // :300-311
}  // namespace g1pubs
namespace g2pubs {
G1Proj sign(const uint8_t *msg, size_t len, const Scalar &sk);                           // g2pubs/bls.go:132-135
G2Proj priv_to_pub(const Scalar &sk);                                                    // :138-140
bool verify(const uint8_t *msg, size_t len, const G2Proj &pub, const G1Proj &sig);       // :159-162
G1Proj aggregate_signatures(const std::vector<G1Proj> &sigs);                            // :165-177
G2Proj aggregate_public_keys(const std::vector<G2Proj> &pubs);                           // :180-186
bool verify_aggregate(const G1Proj &sig, const std::vector<G2Proj> &pubs,
                      const std::vector<std::string> &msgs);                             // :240-270
bool verify_aggregate_common(const G1Proj &sig, const std::vector<G2Proj> &pubs,
                             const uint8_t *msg, size_t len);                            // :275-278
}  // namespace g2pubs

}  // namespace orc
