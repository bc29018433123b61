// oracle_capi.cc -- flat C entry points over the CPU restatement, for ctypes.
// TEST INFRASTRUCTURE ONLY (see oracle.hpp).  Buffers use the same POD layout as
// include/b381.h: fp = 6 x u64 LE limbs (Montgomery, canonical); fp2 = c0||c1;
// fp12 = c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2; g1 affine = x, y, u8 inf + 7 pad
// (104 B); g1 jacobian = x, y, z (144 B); g2 affine 200 B; g2 jacobian 288 B;
// scalar = 4 x u64 LE canonical.
#include "oracle.hpp"
#include <thread>
#include <chrono>

using namespace orc;

namespace {
struct PG1A { Fq x, y; uint8_t inf; uint8_t pad[7]; };
struct PG2A { Fq2 x, y; uint8_t inf; uint8_t pad[7]; };
static_assert(sizeof(Fq) == 48 && sizeof(Fq2) == 96 && sizeof(Fq6) == 288 && sizeof(Fq12) == 576, "pod");
static_assert(sizeof(PG1A) == 104 && sizeof(PG2A) == 200, "pod");
static_assert(sizeof(G1Proj) == 144 && sizeof(G2Proj) == 288, "pod");

inline G1Affine ld(const PG1A &p) { return G1Affine{p.x, p.y, p.inf != 0}; }
inline G2Affine ld(const PG2A &p) { return G2Affine{p.x, p.y, p.inf != 0}; }
inline void st(PG1A &o, const G1Affine &a) { memset(&o, 0, sizeof o); o.x = a.x; o.y = a.y; o.inf = a.infinity; }
inline void st(PG2A &o, const G2Affine &a) { memset(&o, 0, sizeof o); o.x = a.x; o.y = a.y; o.inf = a.infinity; }

template <class FN> void parallel_for(size_t n, int threads, FN fn) {
    if (threads <= 1 || n < 2) { for (size_t i = 0; i < n; i++) fn(i); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
        th.emplace_back([=] { for (size_t i = t; i < n; i += threads) fn(i); });
    for (auto &t : th) t.join();
}
}  // namespace

extern "C" {

uint64_t orc_fq_mul_count() { return g_fq_mul_count; }
void orc_fq_mul_count_reset() { g_fq_mul_count = 0; }

// L0 primitives (stub_fallback.go) -- for the primitivefuncs_test.go golden vectors
void orc_multiply_fq_repr(const u64 a[6], const u64 b[6], u64 hi[6], u64 lo[6]) { multiply_fq_repr(a, b, hi, lo); }
void orc_mont_reduce(const u64 hi[6], const u64 lo[6], u64 out[6]) { Repr r = mont_reduce(hi, lo); memcpy(out, r.l, 48); }
u64 orc_mac_with_carry(u64 a, u64 b, u64 c, u64 *carry) { return mac_with_carry(a, b, c, *carry); }
u64 orc_add_with_carry(u64 a, u64 b, u64 *carry) { return add_with_carry(a, b, *carry); }
u64 orc_sub_with_borrow(u64 a, u64 b, u64 *borrow) { return sub_with_borrow(a, b, *borrow); }

// Fq: op 0 mul, 1 add, 2 sub, 3 square, 4 neg, 5 double, 6 inverse (0 -> 0), 7 FQReprToFQ, 8 ToRepr, 9 sqrt (fail -> 0)
void orc_fq_op(int op, const Fq *a, const Fq *b, Fq *out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        Fq x = a[i];
        switch (op) {
            case 0: x.mul(b[i]); break;
            case 1: x.add(b[i]); break;
            case 2: x.sub(b[i]); break;
            case 3: x.square(); break;
            case 4: x.neg(); break;
            case 5: x.dbl(); break;
            case 6: { Fq r; if (x.inverse(r)) x = r; else x = FQ_ZERO; break; }
            case 7: x = fq_from_repr(x.n); break;
            case 8: x.n = x.to_repr(); break;
            case 9: { Fq r; if (x.sqrt(r)) x = r; else x = FQ_ZERO; break; }
        }
        out[i] = x;
    }
}
// Fq2: op 0 mul, 1 add, 2 sub, 3 square, 4 neg, 5 double, 6 inverse, 7 frobenius(1), 8 mul_by_nonresidue, 9 sqrt
void orc_fq2_op(int op, const Fq2 *a, const Fq2 *b, Fq2 *out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        Fq2 x = a[i];
        switch (op) {
            case 0: x.mul(b[i]); break;
            case 1: x.add(b[i]); break;
            case 2: x.sub(b[i]); break;
            case 3: x.square(); break;
            case 4: x.neg(); break;
            case 5: x.dbl(); break;
            case 6: if (!x.inverse()) x = FQ2_ZERO; break;
            case 7: x.frobenius(1); break;
            case 8: x.mul_by_nonresidue(); break;
            case 9: { Fq2 r; if (x.sqrt(r)) x = r; else x = FQ2_ZERO; break; }
        }
        out[i] = x;
    }
}
// Fq6: op 0 mul, 1 add, 2 sub, 3 square, 4 neg, 6 inverse, 7 frobenius(arg), 8 mul_by_nonresidue,
//      10 mul_by_1 (b[i].c1), 11 mul_by_01 (b[i].c0, b[i].c1)
void orc_fq6_op(int op, int arg, const Fq6 *a, const Fq6 *b, Fq6 *out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        Fq6 x = a[i];
        switch (op) {
            case 0: x.mul(b[i]); break;
            case 1: x.add(b[i]); break;
            case 2: x.sub(b[i]); break;
            case 3: x.square(); break;
            case 4: x.neg(); break;
            case 6: if (!x.inverse()) x = FQ6_ZERO; break;
            case 7: x.frobenius(arg); break;
            case 8: x.mul_by_nonresidue(); break;
            case 10: x.mul_by_1(b[i].c1); break;
            case 11: x.mul_by_01(b[i].c0, b[i].c1); break;
        }
        out[i] = x;
    }
}
// Fq12: op 0 mul, 3 square, 6 inverse (0 -> 0), 7 frobenius(arg), 12 conjugate,
//       13 mul_by_014 (b[i].c0.c0, b[i].c0.c1, b[i].c1.c1), 14 exp by u64 arg64 (FQ12.Exp)
void orc_fq12_op(int op, uint64_t arg, const Fq12 *a, const Fq12 *b, Fq12 *out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        Fq12 x = a[i];
        switch (op) {
            case 0: x.mul(b[i]); break;
            case 3: x.square(); break;
            case 6: if (!x.inverse()) { x.c0 = FQ6_ZERO; x.c1 = FQ6_ZERO; } break;
            case 7: x.frobenius((unsigned)arg); break;
            case 12: x.conjugate(); break;
            case 13: x.mul_by_014(b[i].c0.c0, b[i].c0.c1, b[i].c1.c1); break;
            case 14: x = x.exp(repr_u64(arg)); break;
        }
        out[i] = x;
    }
}

// ---- constants -------------------------------------------------------------
void orc_g1_generator(PG1A *o) { st(*o, G1_AFFINE_ONE); }
void orc_g2_generator(PG2A *o) { st(*o, G2_AFFINE_ONE); }
void orc_fq_one(Fq *o) { *o = FQ_ONE; }
void orc_frobenius_tables(Fq2 *fq6c1, Fq2 *fq6c2, Fq2 *fq12c1);

// ---- G1 --------------------------------------------------------------------
void orc_g1_add(const G1Proj *a, const G1Proj *b, G1Proj *o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = g1_add(a[i], b[i]); }
void orc_g1_add_affine(const G1Proj *a, const PG1A *b, G1Proj *o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = g1_add_affine(a[i], ld(b[i])); }
void orc_g1_double(const G1Proj *a, G1Proj *o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = g1_double(a[i]); }
void orc_g1_to_affine(const G1Proj *a, PG1A *o, size_t n) { for (size_t i = 0; i < n; i++) st(o[i], g1_to_affine(a[i])); }
void orc_g1_to_proj(const PG1A *a, G1Proj *o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = g1_to_proj(ld(a[i])); }
void orc_g1_affine_mul_fr(const PG1A *a, const Scalar *s, G1Proj *o, size_t n, int threads) {
    parallel_for(n, threads, [=](size_t i) { o[i] = g1_affine_mul_fr(ld(a[i]), s[i]); });
}
// AggregatePublicKeys, g1pubs/bls.go:192-198: left fold of G1Projective.Add starting from zero
void orc_g1_sum_proj(const G1Proj *a, size_t n, G1Proj *o) {
    G1Proj acc = G1_PROJ_ZERO;
    for (size_t i = 0; i < n; i++) acc = g1_add(acc, a[i]);
    *o = acc;
}
void orc_g1_sum_affine(const PG1A *a, size_t n, G1Proj *o) {
    G1Proj acc = G1_PROJ_ZERO;
    for (size_t i = 0; i < n; i++) acc = g1_add(acc, g1_to_proj(ld(a[i])));
    *o = acc;
}
// sum_i s_i * P_i with the reference's double-and-add (G1Affine.MulFR) and Add fold
void orc_g1_msm_naive(const PG1A *a, const Scalar *s, size_t n, G1Proj *o, int threads) {
    std::vector<G1Proj> part(n);
    parallel_for(n, threads, [&](size_t i) { part[i] = g1_affine_mul_fr(ld(a[i]), s[i]); });
    G1Proj acc = G1_PROJ_ZERO;
    for (size_t i = 0; i < n; i++) acc = g1_add(acc, part[i]);
    *o = acc;
}
int orc_g1_proj_equal(const G1Proj *a, const G1Proj *b) { return g1_proj_equal(*a, *b); }
int orc_g1_is_on_curve(const PG1A *a) { return g1_is_on_curve(ld(*a)); }
int orc_g1_in_subgroup(const PG1A *a) { return g1_in_subgroup(ld(*a)); }
void orc_g1_compress(const PG1A *a, uint8_t out[48]) { g1_compress(ld(*a), out); }
int orc_g1_decompress(const uint8_t in[48], PG1A *o, int checked) {
    G1Affine a = G1_AFFINE_ZERO;
    int e = checked ? g1_decompress(in, a) : g1_decompress_unchecked(in, a);
    st(*o, a);
    return e;
}

// ---- G2 --------------------------------------------------------------------
void orc_g2_add(const G2Proj *a, const G2Proj *b, G2Proj *o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = g2_add(a[i], b[i]); }
void orc_g2_add_affine(const G2Proj *a, const PG2A *b, G2Proj *o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = g2_add_affine(a[i], ld(b[i])); }
void orc_g2_double(const G2Proj *a, G2Proj *o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = g2_double(a[i]); }
void orc_g2_to_affine(const G2Proj *a, PG2A *o, size_t n) { for (size_t i = 0; i < n; i++) st(o[i], g2_to_affine(a[i])); }
void orc_g2_to_proj(const PG2A *a, G2Proj *o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = g2_to_proj(ld(a[i])); }
void orc_g2_affine_mul_fr(const PG2A *a, const Scalar *s, G2Proj *o, size_t n, int threads) {
    parallel_for(n, threads, [=](size_t i) { o[i] = g2_affine_mul_fr(ld(a[i]), s[i]); });
}
void orc_g2_sum_proj(const G2Proj *a, size_t n, G2Proj *o) {  // AggregateSignatures, g1pubs/bls.go:177-183
    G2Proj acc = G2_PROJ_ZERO;
    for (size_t i = 0; i < n; i++) acc = g2_add(acc, a[i]);
    *o = acc;
}
void orc_g2_sum_affine(const PG2A *a, size_t n, G2Proj *o) {
    G2Proj acc = G2_PROJ_ZERO;
    for (size_t i = 0; i < n; i++) acc = g2_add(acc, g2_to_proj(ld(a[i])));
    *o = acc;
}
int orc_g2_proj_equal(const G2Proj *a, const G2Proj *b) { return g2_proj_equal(*a, *b); }
int orc_g2_is_on_curve(const PG2A *a) { return g2_is_on_curve(ld(*a)); }
int orc_g2_in_subgroup(const PG2A *a) { return g2_in_subgroup(ld(*a)); }
void orc_g2_compress(const PG2A *a, uint8_t out[96]) { g2_compress(ld(*a), out); }
int orc_g2_decompress(const uint8_t in[96], PG2A *o, int checked) {
    G2Affine a = G2_AFFINE_ZERO;
    int e = checked ? g2_decompress(in, a) : g2_decompress_unchecked(in, a);
    st(*o, a);
    return e;
}

// ---- pairing ---------------------------------------------------------------
// coefficient table of G2AffineToPrepared: returns the number of triples (68), writes 3*68 fp2
int orc_g2_prepare(const PG2A *q, Fq2 *coeffs) {
    G2Prepared p = g2_prepare(ld(*q));
    if (coeffs) memcpy(coeffs, p.coeffs.data(), p.coeffs.size() * sizeof(Fq2));
    return (int)p.nsteps();
}
// MillerLoop over npairs (P_i, Q_i): one Fq12 product
void orc_miller_loop(const PG1A *p, const PG2A *q, size_t npairs, Fq12 *out) {
    std::vector<G2Prepared> prep(npairs);
    std::vector<MillerItem> items(npairs);
    for (size_t i = 0; i < npairs; i++) { prep[i] = g2_prepare(ld(q[i])); items[i] = MillerItem{ld(p[i]), &prep[i]}; }
    *out = miller_loop(items);
}
int orc_final_exp(const Fq12 *in, Fq12 *out) { return final_exponentiation(*in, *out) ? 1 : 0; }
// n independent bls.Pairing calls on affine inputs (z = 1 Jacobian inputs of the reference)
void orc_pairing_batch(const PG1A *p, const PG2A *q, size_t n, Fq12 *out, int threads) {
    parallel_for(n, threads, [=](size_t i) { out[i] = pairing(g1_to_proj(ld(p[i])), g2_to_proj(ld(q[i]))); });
}
// generalised CompareTwoPairings: for group g, FE(prod_{i in [off[g],off[g+1])} ML(P_i,Q_i)) == 1
void orc_pairing_product_is_one(const PG1A *p, const PG2A *q, const uint32_t *off, size_t ngroups, uint8_t *ok, int threads) {
    parallel_for(ngroups, threads, [=](size_t g) {
        Fq12 f, r;
        orc_miller_loop(p + off[g], q + off[g], off[g + 1] - off[g], &f);
        ok[g] = (final_exponentiation(f, r) && r == FQ12_ONE) ? 1 : 0;
    });
}
int orc_compare_two_pairings(const G1Proj *p1, const G2Proj *q1, const G1Proj *p2, const G2Proj *q2) {
    return compare_two_pairings(*p1, *q1, *p2, *q2);
}

// ---- RNG / scalars -----------------------------------------------------------
void orc_rand_fr(uint64_t *state, Scalar *out, size_t n) {
    XorShift r(*state);
    for (size_t i = 0; i < n; i++) out[i] = rand_fr(r);
    *state = r.state;
}
void orc_rand_fq(uint64_t *state, Fq *out, size_t n) {
    XorShift r(*state);
    for (size_t i = 0; i < n; i++) out[i] = rand_fq(r);
    *state = r.state;
}
void orc_scalar_mul_mod_r(const Scalar *a, const Scalar *b, Scalar *o) { *o = scalar_mul_mod_r(*a, *b); }
void orc_scalar_add_mod_r(const Scalar *a, const Scalar *b, Scalar *o) { *o = scalar_add_mod_r(*a, *b); }

// ---- timing helpers for bench.py's cpu_baseline / --impl reference ------------
// Runs n pairings (inputs cycled from the m given pairs) on `threads` host threads, returns seconds.
double orc_time_pairings(const PG1A *p, const PG2A *q, size_t m, size_t n, int threads, Fq12 *last) {
    std::vector<Fq12> outs(threads > 0 ? threads : 1);
    auto t0 = std::chrono::steady_clock::now();
    parallel_for(n, threads, [&](size_t i) {
        outs[threads > 1 ? i % threads : 0] = pairing(g1_to_proj(ld(p[i % m])), g2_to_proj(ld(q[i % m])));
    });
    auto t1 = std::chrono::steady_clock::now();
    if (last) *last = outs[0];
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"

namespace orc { const Fq2 *frob_fq6_c1(); const Fq2 *frob_fq6_c2(); const Fq2 *frob_fq12_c1(); }
extern "C" void orc_frobenius_tables(Fq2 *fq6c1, Fq2 *fq6c2, Fq2 *fq12c1) {
    memcpy(fq6c1, orc::frob_fq6_c1(), 6 * sizeof(Fq2));
    memcpy(fq6c2, orc::frob_fq6_c2(), 6 * sizeof(Fq2));
    memcpy(fq12c1, orc::frob_fq12_c1(), 12 * sizeof(Fq2));
}
