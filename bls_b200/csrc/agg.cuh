// agg.cuh -- aggregation kernels: point sums (AggregatePublicKeys / AggregateSignatures,
// g1pubs/bls.go:177-204), the Pippenger bucket MSM the north star adds for weighted aggregates,
// and the per-attestation public-key aggregation of the VerifyAggregateCommon batch
// (g1pubs/bls.go:287-290).  The reference folds serially with G1Projective.Add (g1.go:400-482);
// here every stage is a data-parallel pass over HBM-resident arrays:
//
//   sum      : grid-stride mixed additions per thread -> shared-memory tree per block -> one block
//   MSM      : digit histogram -> per-window exclusive scan -> scatter of point indices by bucket
//              (a counting sort; 4 B per point per window) -> fixed-size CHUNK partial sums (one
//              thread per chunk, so a bucket holding every point costs the same as a uniform
//              spread) -> in-bucket chunk tree -> running-sum segments -> window sums -> Horner
//
// Only canonical results (normalised points) cross the ABI, so the order of additions -- which
// differs from the reference's left fold -- is not observable.
#pragma once
#include "curve.cuh"

namespace b381 {

#if defined(__CUDACC__)

#define MSM_CHUNK 64u       // points per chunk partial sum
#define MSM_SEG 16u         // buckets per running-sum segment (the padding of the bucket count; msm_geom::seg is the length in use)
#define MSM_SEG_SMALL 4u    // ... of a small MSM (MSM_SMALL_CHUNKS): its bucket reduction is one dependent chain per segment, shorter segments shorten it
#define MSM_LANE_REDUCE_MAX_SEGS 8192u   // windows x segments on this rank up to which the G1 bucket reduction runs four lanes per segment
#define MSM_SMALL_CHUNKS 32768u // an MSM with at most this many chunk partials in all takes the tree rounds beyond MSM_FOLD_SMALL partials per bucket:
#define MSM_FOLD_SMALL 4u       // its rounds are small launches, while a serial fold of a few big buckets (group-by sums, the short top window of 64-bit weights) is one long chain
#define MSM_FOLD_MAX 64u    // up to this many chunk partials per bucket (uniform scalars: 2-5) are added by one thread per bucket (k_msm_bucket_fold);
                            // beyond that (skewed scalars, up to every point in one bucket) the in-bucket tree rounds run first

// ---- block tree over shared memory: the sum of every thread's acc ends up in thread 0 -------------
template <class F, int BLOCK> __device__ void block_reduce_xyzz(xyzz<F> &acc, xyzz<F> *sm) {
    sm[threadIdx.x] = acc;
    __syncthreads();
#pragma unroll 1
    for (int s = BLOCK / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            xyzz<F> a = sm[threadIdx.x];
            xyzz_add(a, sm[threadIdx.x + s]);
            sm[threadIdx.x] = a;
        }
        __syncthreads();
    }
    acc = sm[0];
}

// partial[b] = sum of the points this block strides over
template <class F, class APOD, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_sum_partial(const APOD *__restrict__ p, size_t n, xyzz<F> *__restrict__ partial) {
    __shared__ xyzz<F> sm[BLOCK];
    xyzz<F> acc;
    xyzz_set_inf(acc);
    size_t stride = (size_t)gridDim.x * BLOCK;
    for (size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x; i < n; i += stride) {
        typename F::T x, y; bool inf;
        load_affine(x, y, inf, p + i);
        if (!inf) xyzz_madd(acc, x, y);
    }
    block_reduce_xyzz<F, BLOCK>(acc, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// out = normalised sum of m partials (one block)
template <class F, class JPOD, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_sum_final(const xyzz<F> *__restrict__ partial, int m, JPOD *__restrict__ out) {
    __shared__ xyzz<F> sm[BLOCK];
    xyzz<F> acc;
    xyzz_set_inf(acc);
    for (int i = threadIdx.x; i < m; i += BLOCK) { xyzz<F> q = partial[i]; xyzz_add(acc, q); }
    block_reduce_xyzz<F, BLOCK>(acc, sm);
    if (threadIdx.x == 0) {
        typename F::T ox, oy, oz;
        xyzz_to_jac_normalised(ox, oy, oz, acc);
        store_jac(out, ox, oy, oz);
    }
}

// out = normalised sum of n Jacobian points (the all-gathered per-rank MSM partials)
__global__ void __launch_bounds__(128) k_g1_fold(const g1_jac_pod *__restrict__ parts, size_t n, g1_jac_pod *__restrict__ out) {
    __shared__ xyzz<FpInl> sm[128];
    xyzz<FpInl> acc;
    xyzz_set_inf(acc);
    for (size_t i = threadIdx.x; i < n; i += 128) {
        fp x, y, z;
        load_jac(x, y, z, parts + i);
        xyzz<FpInl> q;
        xyzz_from_jac(q, x, y, z);
        xyzz_add(acc, q);
    }
    block_reduce_xyzz<FpInl, 128>(acc, sm);
    if (threadIdx.x == 0) {
        fp ox, oy, oz;
        xyzz_to_jac_normalised(ox, oy, oz, acc);
        store_jac(out, ox, oy, oz);
    }
}

// ---- Pippenger MSM over G1 ------------------------------------------------------------------------
// windows handled by a call: w = w0 + j * wstep, j < nw (wstep = number of ranks when bucket-sharded)
struct msm_geom { int c, w0, wstep, nw; uint32_t nb; uint32_t maxchunks; uint32_t seg; size_t n; };   // seg: buckets per running-sum segment

__global__ void k_msm_hist(const uint64_t *__restrict__ k, msm_geom g, uint32_t *__restrict__ count) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    uint64_t s[4] = {k[4 * i], k[4 * i + 1], k[4 * i + 2], k[4 * i + 3]};
    // signed digits, carry chain walked once over all windows up to this rank's last one
    uint32_t carry = 0;
    const uint32_t half = 1u << (g.c - 1);
    int j = 0;
    for (int w = 0; j < g.nw; w++) {
        uint32_t raw = msm_digit(s, w, g.c) + carry, mag;
        if (raw > half) { mag = (1u << g.c) - raw; carry = 1; } else { mag = raw; carry = 0; }
        if (w == g.w0 + j * g.wstep) {
            if (mag) atomicAdd(&count[(size_t)j * g.nb + mag], 1u);
            j++;
        }
    }
}

// per window (one block): bucket_off = exclusive scan of count; the sorted index list of the window is cut into RUNS of
// MSM_CHUNK consecutive positions (one per thread of k_msm_chunk_sum, whatever buckets they fall in), so bucket d owns
// one chunk partial sum per run it touches: nch(d) = last run - first run + 1.  chunk_off = exclusive scan of nch; both get a
// closing total at [nb].  maxch = largest chunk count of a bucket.
__global__ void __launch_bounds__(1024) k_msm_scan(const uint32_t *__restrict__ count, msm_geom g, uint32_t *__restrict__ bucket_off,
                                                   uint32_t *__restrict__ chunk_off, uint32_t *__restrict__ maxch) {
    __shared__ uint32_t sa[1024];
    int j = blockIdx.x, t = threadIdx.x;
    const uint32_t *cnt = count + (size_t)j * g.nb;
    uint32_t *bo = bucket_off + (size_t)j * (g.nb + 1), *co = chunk_off + (size_t)j * (g.nb + 1);
    uint32_t per = (g.nb + 1023u) / 1024u;
    uint32_t lo = t * per, hi = lo + per < g.nb ? lo + per : g.nb;
    if (lo > g.nb) lo = g.nb;
    uint32_t a = 0, b = 0, mx = 0;
    for (uint32_t d = lo; d < hi; d++) a += cnt[d];
    sa[t] = a;
    __syncthreads();
    for (int s = 1; s < 1024; s <<= 1) {          // inclusive Hillis-Steele scan
        uint32_t va = t >= s ? sa[t - s] : 0;
        __syncthreads();
        sa[t] += va;
        __syncthreads();
    }
    uint32_t ea = sa[t] - a, tot = sa[1023];
    __syncthreads();
    for (uint32_t d = lo; d < hi; d++) {
        uint32_t c = cnt[d];
        bo[d] = ea;
        uint32_t ch = c ? ((ea + c - 1) / MSM_CHUNK - ea / MSM_CHUNK + 1) : 0;
        b += ch; mx = ch > mx ? ch : mx;
        ea += c;
    }
    sa[t] = b;
    __syncthreads();
    for (int s = 1; s < 1024; s <<= 1) {
        uint32_t vb = t >= s ? sa[t - s] : 0;
        __syncthreads();
        sa[t] += vb;
        __syncthreads();
    }
    uint32_t eb = sa[t] - b;
    for (uint32_t d = lo; d < hi; d++) {
        uint32_t c = cnt[d], e0 = bo[d];
        co[d] = eb;
        eb += c ? ((e0 + c - 1) / MSM_CHUNK - e0 / MSM_CHUNK + 1) : 0;
    }
    if (t == 1023) { bo[g.nb] = tot; co[g.nb] = sa[t]; }
    if (mx) atomicMax(maxch, mx);
}

// idx[j][bucket_off[j][d] ..] = the indices of the points whose digit in window j is d (count is consumed)
__global__ void k_msm_scatter(const uint64_t *__restrict__ k, msm_geom g, const uint32_t *__restrict__ bucket_off,
                              uint32_t *__restrict__ count, uint32_t *__restrict__ idx) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    uint64_t s[4] = {k[4 * i], k[4 * i + 1], k[4 * i + 2], k[4 * i + 3]};
    uint32_t carry = 0;
    const uint32_t half = 1u << (g.c - 1);
    int j = 0;
    for (int w = 0; j < g.nw; w++) {
        uint32_t raw = msm_digit(s, w, g.c) + carry, mag, neg;
        if (raw > half) { mag = (1u << g.c) - raw; carry = 1; neg = MSM_NEG_BIT; } else { mag = raw; carry = 0; neg = 0; }
        if (w == g.w0 + j * g.wstep) {
            if (mag) {
                uint32_t slot = atomicSub(&count[(size_t)j * g.nb + mag], 1u) - 1u;
                idx[(size_t)j * g.n + bucket_off[(size_t)j * (g.nb + 1) + mag] + slot] = (uint32_t)i | neg;
            }
            j++;
        }
    }
}

// one thread per RUN of MSM_CHUNK consecutive positions of the window's sorted index list: every lane of a warp performs the
// same number of mixed additions whatever the bucket sizes are (with one thread per bucket-aligned chunk a third of the lanes
// idled on the short tail chunks of buckets holding a little more than MSM_CHUNK points: ncu, profiles/r02_experiments.md).
// At a bucket boundary the accumulator is flushed as that bucket's chunk number (run - first run of the bucket).
template <class F, class APOD> __global__ void __launch_bounds__(128) k_msm_chunk_sum(const APOD *__restrict__ pts, const uint32_t *__restrict__ idx, msm_geom g,
                                                       const uint32_t *__restrict__ bucket_off, const uint32_t *__restrict__ chunk_off,
                                                       xyzz<F> *__restrict__ chunks, uint32_t *__restrict__ chunk_bucket) {
    int j = blockIdx.y;
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t *co = chunk_off + (size_t)j * (g.nb + 1), *bo = bucket_off + (size_t)j * (g.nb + 1);
    const uint32_t total = bo[g.nb];
    if ((uint64_t)q * MSM_CHUNK >= total) return;
    const uint32_t first = q * MSM_CHUNK, last = first + MSM_CHUNK < total ? first + MSM_CHUNK : total;
    const uint32_t *ix = idx + (size_t)j * g.n;
    xyzz<F> *cbase = chunks + (size_t)j * g.maxchunks;
    uint32_t *bbase = chunk_bucket + (size_t)j * g.maxchunks;
    uint32_t b = 0, bend = 0;
    typedef typename acc_policy<F>::type FA;           // same element type; G1: the two squarings of a mixed addition use fp_sqr_v
    xyzz<FA> accv;
    xyzz<F> &acc = reinterpret_cast<xyzz<F> &>(accv);
    xyzz_set_inf(accv);
#pragma unroll 1
    for (uint32_t pos = first; pos < last; pos++) {
        if (pos == bend || pos == first) {
            if (pos != first) { uint32_t ci = co[b] + (q - bo[b] / MSM_CHUNK); cbase[ci] = acc; bbase[ci] = b; xyzz_set_inf(accv); }
            uint32_t lo = 0, hi = g.nb;            // the bucket of position pos: largest b with bo[b] <= pos (it is non-empty)
            while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (bo[mid] <= pos) lo = mid; else hi = mid; }
            b = lo; bend = bo[b + 1];
        }
        typename F::T x, y; bool inf;
        const uint32_t e = ix[pos];
        load_affine(x, y, inf, pts + (e & ~MSM_NEG_BIT));
        if (e & MSM_NEG_BIT) F::neg(y, y);             // a negative digit adds -P
        if (!inf) xyzz_madd(accv, x, y);
    }
    uint32_t ci = co[b] + (q - bo[b] / MSM_CHUNK);
    cbase[ci] = acc; bbase[ci] = b;
}

// round r of the in-bucket tree: chunk l of a bucket absorbs chunk l + 2^r when l % 2^(r+1) == 0.  Grid-stride over the chunks:
// uniform scalars need two rounds, the launches of the remaining ones (a single bucket may hold every point) exit at once.
template <class F> __global__ void __launch_bounds__(128) k_msm_chunk_tree(xyzz<F> *__restrict__ chunks, const uint32_t *__restrict__ chunk_bucket,
                                                        const uint32_t *__restrict__ chunk_off, msm_geom g, int r,
                                                        const uint32_t *__restrict__ maxch, uint32_t fold_max) {
    if ((1u << r) >= *maxch || *maxch <= fold_max) return;               // few chunks per bucket: k_msm_bucket_fold adds them
    int j = blockIdx.y;
    const uint32_t *co = chunk_off + (size_t)j * (g.nb + 1);
    const uint32_t nchunks = co[g.nb];
    xyzz<F> *base = chunks + (size_t)j * g.maxchunks;
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < nchunks; q += gridDim.x * blockDim.x) {
        uint32_t b = chunk_bucket[(size_t)j * g.maxchunks + q];
        uint32_t l = q - co[b], nch = co[b + 1] - co[b];
        if ((l & ((2u << r) - 1)) || l + (1u << r) >= nch) continue;
        xyzz<F> a = base[q];
        xyzz_add(a, base[q + (1u << r)]);
        base[q] = a;
    }
}

// one thread per (window, bucket): chunk 0 of the bucket absorbs its other partials.  With one thread per chunk and tree rounds
// only every second / fourth lane works and every round is a launch over all chunks (4.0 ms at 2^22 points); here all lanes of a
// warp fold a similar, small number of partials.
template <class F> __global__ void __launch_bounds__(128) k_msm_bucket_fold(xyzz<F> *__restrict__ chunks, const uint32_t *__restrict__ chunk_off, msm_geom g,
                                                         const uint32_t *__restrict__ maxch, uint32_t fold_max) {
    if (*maxch > fold_max || *maxch < 2) return;                          // the tree rounds ran (or there is nothing to add)
    int j = blockIdx.y;
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= g.nb) return;
    const uint32_t *co = chunk_off + (size_t)j * (g.nb + 1);
    const uint32_t c0 = co[b], c1 = co[b + 1];
    if (c1 - c0 < 2) return;
    xyzz<F> *base = chunks + (size_t)j * g.maxchunks;
    xyzz<F> a = base[c0];
#pragma unroll 1
    for (uint32_t q = c0 + 1; q < c1; q++) { xyzz<F> t = base[q]; xyzz_add(a, t); }
    base[c0] = a;
}

// ---- G1 group law on a group of FOUR LANES holding the same operands ------------------------------------------------------
// The reductions behind the chunk sums (bucket running sums, window sums, window shifts) are chains of dependent additions with
// little parallel work: a thread advances one Fq multiplication per ~1 us, an XYZZ addition is 14 of them.  Four lanes run the
// independent multiplications of an addition / doubling side by side (4 / 3 levels instead of 14 / 9) and exchange the results
// by SHFL; every lane of the group ends with the same value, bit-identical to xyzz_add / xyzz_dbl (same formulas).
struct lane_grp { unsigned sub, base, mask; };
__device__ __forceinline__ lane_grp lane_group() {
    const unsigned lane = threadIdx.x & 31u;
    lane_grp g; g.sub = lane & 3u; g.base = lane & ~3u; g.mask = 0xFu << g.base; return g;
}
__device__ __forceinline__ void lane_pick(fp &r, unsigned sub, const fp &a0, const fp &a1, const fp &a2, const fp &a3) {
#pragma unroll
    for (int k = 0; k < 12; k++) r.l[k] = sub == 0 ? a0.l[k] : (sub == 1 ? a1.l[k] : (sub == 2 ? a2.l[k] : a3.l[k]));
}
__device__ __forceinline__ void lane_bcast(fp &r, const fp &v, const lane_grp &g, unsigned src) {
#pragma unroll
    for (int k = 0; k < 12; k++) r.l[k] = __shfl_sync(g.mask, v.l[k], g.base + src);
}
// p = 2p (dbl-2008-s-1 as xyzz_dbl_body): V = U^2 | M = X^2  ->  W = U V | S = X V | MM = (3M)^2 | ZZ3 = V ZZ  ->
// W Y | 3M (S - X3) | W ZZZ
__device__ __noinline__ void xyzz_dbl_lanes(xyzz<FpInl> &p) {
    if (xyzz_is_inf(p)) return;
    const lane_grp g = lane_group();
    fp u, a, b, r, v, m, w, s, mm, zz3, t, wy, zzz3;
    fp_add(u, p.y, p.y);
    lane_pick(a, g.sub, u, p.x, u, p.x);
    r = fp_mul_v(a, a);
    lane_bcast(v, r, g, 0); lane_bcast(m, r, g, 1);
    fp_add(t, m, m); fp_add(m, t, m);                  // 3 X^2
    lane_pick(a, g.sub, u, p.x, m, v); lane_pick(b, g.sub, v, v, m, p.zz);
    r = fp_mul_v(a, b);
    lane_bcast(w, r, g, 0); lane_bcast(s, r, g, 1); lane_bcast(mm, r, g, 2); lane_bcast(zz3, r, g, 3);
    fp_sub(mm, mm, s); fp_sub(mm, mm, s);              // X3
    fp_sub(t, s, mm);
    lane_pick(a, g.sub, w, m, w, w); lane_pick(b, g.sub, p.y, t, p.zzz, p.zzz);
    r = fp_mul_v(a, b);
    lane_bcast(wy, r, g, 0); lane_bcast(t, r, g, 1); lane_bcast(zzz3, r, g, 2);
    p.x = mm; fp_sub(p.y, t, wy); p.zz = zz3; p.zzz = zzz3;
}
// p += q (add-2008-s as xyzz_add_body, same case split): U1 | U2 | S1 | S2  ->  PP | RR | ZZ1 ZZ2 | ZZZ1 ZZZ2  ->
// PPP | Q | ZZ3  ->  R (Q - X3) | S1 PPP | ZZZ3
__device__ __noinline__ void xyzz_add_lanes(xyzz<FpInl> &p, const xyzz<FpInl> &q) {
    if (xyzz_is_inf(q)) return;
    if (xyzz_is_inf(p)) { p = q; return; }
    const lane_grp g = lane_group();
    fp a, b, r, u1, u2, s1, s2, pp, rr, zz12, zzz12, ppp, qq, zz3, t, ya, yb, zzz3;
    lane_pick(a, g.sub, p.x, q.x, p.y, q.y); lane_pick(b, g.sub, q.zz, p.zz, q.zzz, p.zzz);
    r = fp_mul_v(a, b);
    lane_bcast(u1, r, g, 0); lane_bcast(u2, r, g, 1); lane_bcast(s1, r, g, 2); lane_bcast(s2, r, g, 3);
    fp_sub(u2, u2, u1);                                // P
    fp_sub(s2, s2, s1);                                // R
    if (fp_is_zero(u2)) {                              // the same on every lane of the group
        if (fp_is_zero(s2)) xyzz_dbl_lanes(p); else xyzz_set_inf(p);
        return;
    }
    lane_pick(a, g.sub, u2, s2, p.zz, p.zzz); lane_pick(b, g.sub, u2, s2, q.zz, q.zzz);
    r = fp_mul_v(a, b);
    lane_bcast(pp, r, g, 0); lane_bcast(rr, r, g, 1); lane_bcast(zz12, r, g, 2); lane_bcast(zzz12, r, g, 3);
    lane_pick(a, g.sub, u2, u1, zz12, zz12);
    r = fp_mul_v(a, pp);
    lane_bcast(ppp, r, g, 0); lane_bcast(qq, r, g, 1); lane_bcast(zz3, r, g, 2);
    fp_sub(t, rr, ppp); fp_sub(t, t, qq); fp_sub(t, t, qq);   // X3 = R^2 - PPP - 2Q
    p.x = t;
    fp_sub(t, qq, t);
    lane_pick(a, g.sub, s2, s1, zzz12, zzz12); lane_pick(b, g.sub, t, ppp, ppp, ppp);
    r = fp_mul_v(a, b);
    lane_bcast(ya, r, g, 0); lane_bcast(yb, r, g, 1); lane_bcast(zzz3, r, g, 2);
    fp_sub(p.y, ya, yb); p.zz = zz3; p.zzz = zzz3;
}
// p = k p, double-and-add MSB first from the top set bit (the value of xyzz_mul_small)
__device__ void xyzz_mul_small_lanes(xyzz<FpInl> &p, uint32_t k) {
    xyzz<FpInl> acc;
    xyzz_set_inf(acc);
#pragma unroll 1
    for (int b = 31 - __clz(k | 1u); b >= 0; b--) {
        xyzz_dbl_lanes(acc);
        if ((k >> b) & 1) xyzz_add_lanes(acc, p);
    }
    p = acc;
}

// the G1 reductions below run on four lanes per unit of work
template <class F> struct lane_shift { static constexpr bool value = false; };
template <> struct lane_shift<FpInl> { static constexpr bool value = true; };

// one thread per (window, segment of MSM_SEG buckets): sum_{d in segment} d * B[d]; chunk 0 of a bucket holds its sum
// LANES (G1 only): FOUR lanes per segment, see above; the grid is sized accordingly by msm_shard_dev.  The lane form shortens the
// chain of a segment by a third but executes twice the instructions, so it is used when the reduction is latency-bound (few
// windows on this rank: a bucket-sharded MSM, or a small one) -- measured at 2^22 points: two windows 1.20 -> 0.90 ms, sixteen
// windows 1.52 -> 2.34 ms.
template <class F, bool LANES> __global__ void __launch_bounds__(128) k_msm_segment_reduce(const xyzz<F> *__restrict__ chunks, const uint32_t *__restrict__ chunk_off,
                                                            msm_geom g, xyzz<F> *__restrict__ segsum) {
    static_assert(!LANES || lane_shift<F>::value, "the lane form exists for G1");
    int j = blockIdx.y;
    uint32_t nseg = g.nb / g.seg;
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) / (LANES ? 4u : 1u);
    if (s >= nseg) return;                             // whole groups leave together (128 is a multiple of 4)
    const uint32_t *co = chunk_off + (size_t)j * (g.nb + 1);
    const xyzz<F> *base = chunks + (size_t)j * g.maxchunks;
    uint32_t lo = s * g.seg, hi = lo + g.seg;
    if (lo == 0) lo = 1;
    xyzz<F> running, acc;
    xyzz_set_inf(running); xyzz_set_inf(acc);
    if constexpr (LANES) {
#pragma unroll 1
        for (uint32_t d = hi; d-- > lo;) {
            const uint32_t c0 = co[d];
            if (co[d + 1] > c0) { xyzz<F> bsum = base[c0]; xyzz_add_lanes(running, bsum); }
            xyzz_add_lanes(acc, running);
        }
        if (lo > 1 && !xyzz_is_inf(running)) { xyzz_mul_small_lanes(running, lo - 1); xyzz_add_lanes(acc, running); }
        if ((threadIdx.x & 3u) == 0) segsum[(size_t)j * nseg + s] = acc;
    } else {
        for (uint32_t d = hi; d-- > lo;) {
            const uint32_t c0 = co[d];
            if (co[d + 1] > c0) { xyzz<F> bsum = base[c0]; xyzz_add(running, bsum); }
            xyzz_add(acc, running);
        }
        if (lo > 1 && !xyzz_is_inf(running)) { xyzz_mul_small(running, lo - 1); xyzz_add(acc, running); }
        segsum[(size_t)j * nseg + s] = acc;
    }
}

// one block per window: winsum[j] = sum of its segment sums
// (LANES: launched with 512 threads = 128 groups of four lanes; otherwise 128 threads)
template <class F, bool LANES> __global__ void __launch_bounds__(LANES ? 512 : 128) k_msm_window_sum(const xyzz<F> *__restrict__ segsum, uint32_t nseg, xyzz<F> *__restrict__ winsum) {
    __shared__ xyzz<F> sm[128];
    int j = blockIdx.x;
    xyzz<F> acc;
    xyzz_set_inf(acc);
    if constexpr (LANES) {
        const unsigned grp = threadIdx.x >> 2, sub = threadIdx.x & 3u;
#pragma unroll 1
        for (uint32_t s = grp; s < nseg; s += 128) { xyzz<F> q = segsum[(size_t)j * nseg + s]; xyzz_add_lanes(acc, q); }
        if (sub == 0) sm[grp] = acc;
        __syncthreads();
#pragma unroll 1
        for (unsigned w = 64; w > 0; w >>= 1) {
            if (grp < w) {                             // whole groups (and, from w = 8 down, part of one warp) take the branch together
                xyzz<F> a = sm[grp], b = sm[grp + w];
                __syncwarp(0xFu << (threadIdx.x & 28u));   // the four lanes hold their copies before lane 0 overwrites the slot
                xyzz_add_lanes(a, b);
                if (sub == 0) sm[grp] = a;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) winsum[j] = sm[0];
    } else {
        for (uint32_t s = threadIdx.x; s < nseg; s += 128) { xyzz<F> q = segsum[(size_t)j * nseg + s]; xyzz_add(acc, q); }
        block_reduce_xyzz<F, 128>(acc, sm);
        if (threadIdx.x == 0) winsum[j] = acc;
    }
}

// The window shift 2^(c w) is a chain of up to 240 doublings, the one serial stretch of the MSM (1.8 ms on one thread: 7.9 us per
// doubling, a quarter of what an 8-GPU shard takes).  A group of FOUR lanes holding the same Jacobian point runs the seven
// multiplications of a doubling (jac_dbl above, same formulas, same values) as three levels:
//   A = X^2 | B = Y^2 | T = Y Z      ->      C = B^2 | S = (X + B)^2 | F = (3A)^2      ->      M = 3A (D - X3)   on every lane
// with the level results exchanged by SHFL (3 x 12 words per level).  Returns with every lane of the group holding the result.
__device__ __noinline__ void jac_dbl_chain_lanes(fp &X, fp &Y, fp &Z, int shifts) {
    const unsigned lane = threadIdx.x & 31u, sub = lane & 3u, base = lane & ~3u, mask = 0xFu << base;
#pragma unroll 1
    for (int i = 0; i < shifts; i++) {
        if (fp_is_zero(Z)) break;                  // infinity stays infinity (the same on every lane of the group)
        fp p, q, r, A, B, T, C, S, Fv, E, D, t;
#pragma unroll
        for (int k = 0; k < 12; k++) {
            p.l[k] = sub == 0 ? X.l[k] : Y.l[k];
            q.l[k] = sub == 0 ? X.l[k] : (sub == 1 ? Y.l[k] : Z.l[k]);
        }
        r = fp_mul_v(p, q);
#pragma unroll
        for (int k = 0; k < 12; k++) {
            A.l[k] = __shfl_sync(mask, r.l[k], base); B.l[k] = __shfl_sync(mask, r.l[k], base + 1); T.l[k] = __shfl_sync(mask, r.l[k], base + 2);
        }
        fp_add(E, A, A); fp_add(E, E, A);          // E = 3A
        fp_add(t, X, B);
#pragma unroll
        for (int k = 0; k < 12; k++) p.l[k] = sub == 0 ? B.l[k] : (sub == 1 ? t.l[k] : E.l[k]);
        r = fp_mul_v(p, p);
#pragma unroll
        for (int k = 0; k < 12; k++) {
            C.l[k] = __shfl_sync(mask, r.l[k], base); S.l[k] = __shfl_sync(mask, r.l[k], base + 1); Fv.l[k] = __shfl_sync(mask, r.l[k], base + 2);
        }
        fp_sub(D, S, A); fp_sub(D, D, C); fp_add(D, D, D);      // D = 2((X + B)^2 - A - C)
        fp_sub(X, Fv, D); fp_sub(X, X, D);                      // X3 = F - 2D
        fp_add(Z, T, T);                                        // Z3 = 2 Y Z
        fp_sub(t, D, X);
        r = fp_mul_v(E, t);
        fp_add(C, C, C); fp_add(C, C, C); fp_add(C, C, C);
        fp_sub(Y, r, C);                                        // Y3 = E (D - X3) - 8C
    }
}

// window j is shifted to its weight 2^(c * w_j) -- G1: by the four lanes 4j .. 4j+3 together, otherwise by thread j --, then a
// tree adds the windows; thread 0 writes the result: normalised (z = 1) for a complete MSM, plain Jacobian for a bucket-sharded partial
template <class F, class JPOD, int BLOCK> __global__ void __launch_bounds__(BLOCK) k_msm_combine(const xyzz<F> *__restrict__ winsum, msm_geom g, int normalise,
                                                    JPOD *__restrict__ out) {
    __shared__ xyzz<F> sm[BLOCK];
    xyzz<F> acc;
    xyzz_set_inf(acc);
    if constexpr (lane_shift<F>::value) {
        const int grp = threadIdx.x >> 2;
        for (int j0 = 0; j0 < g.nw; j0 += BLOCK / 4) {   // uniform over the block: the shuffles of a group need its four lanes
            const int j = j0 + grp;
            const bool have = j < g.nw;
            xyzz<F> s;
            if (have) s = winsum[j]; else xyzz_set_inf(s);
            int shifts = have ? g.c * (g.w0 + j * g.wstep) : 0;
            if (!xyzz_is_inf(s) && shifts) {
                fp x, y, z;
                F::mul(x, s.x, s.zz);              // (X ZZ, Y ZZZ, ZZ) is the same point in Jacobian form
                F::mul(y, s.y, s.zzz);
                z = s.zz;
                jac_dbl_chain_lanes(x, y, z, shifts);
                s.x = x; s.y = y;
                F::sqr(s.zz, z);
                F::mul(s.zzz, s.zz, z);
            }
            if ((threadIdx.x & 3) == 0) xyzz_add(acc, s);
        }
    } else {
    for (int j = threadIdx.x; j < g.nw; j += BLOCK) {
        xyzz<F> s = winsum[j];
        int shifts = g.c * (g.w0 + j * g.wstep);
        if (!xyzz_is_inf(s) && shifts) {           // the shift is a chain of doublings: Jacobian form (2M + 5S each),
            typedef typename shift_policy<F>::type FS;   // multiplications inlined so that independent products overlap
            jac_pt<FS> js;
            F::mul(js.x, s.x, s.zz);               // (X ZZ, Y ZZZ, ZZ) is the same point in Jacobian form
            F::mul(js.y, s.y, s.zzz);
            js.z = s.zz;
#pragma unroll 1
            for (int i = 0; i < shifts; i++) jac_dbl(js);
            s.x = js.x; s.y = js.y;
            F::sqr(s.zz, js.z);
            F::mul(s.zzz, s.zz, js.z);
        }
        xyzz_add(acc, s);
    }
    }
    block_reduce_xyzz<F, BLOCK>(acc, sm);
    if (threadIdx.x == 0) {
        typename F::T ox, oy, oz;
        if (normalise) xyzz_to_jac_normalised(ox, oy, oz, acc);
        else if (xyzz_is_inf(acc)) { F::set_zero(ox); F::set_one(oy); F::set_zero(oz); }
        else {   // XYZZ -> Jacobian with Z = ZZZ/ZZ... avoided: (X*ZZZ^2*.., ) use Z = ZZ: X' = X*ZZ, Y' = Y*ZZZ, Z' = ZZ
            // (X/ZZ, Y/ZZZ) == (X*ZZ / ZZ^2, Y*ZZZ / ZZ^3) because ZZZ^2 == ZZ^3
            F::mul(ox, acc.x, acc.zz);
            F::mul(oy, acc.y, acc.zzz);
            oz = acc.zz;
        }
        store_jac(out, ox, oy, oz);
    }
}

// ---- VerifyAggregateCommon batch: per-attestation public-key aggregation -----------------------------
// attestation a: pk = sum registry[key_idx[key_off[a] .. key_off[a+1])]; writes the two Miller pairs of
// CompareTwoPairings(G1One, sig, pk, H(m)) (g1pubs/bls.go:165-168, pairing.go:140-147):
//   (P, Q)[2a] = (G1One, sig[a]),  (P, Q)[2a+1] = (-pk, msg_hash[msg_idx[a]])
__global__ void __launch_bounds__(128) k_attest_pairs(const g1_affine_pod *__restrict__ registry, const uint32_t *__restrict__ key_idx,
                                                      const uint32_t *__restrict__ key_off, const g2_affine_pod *__restrict__ sig,
                                                      const g2_affine_pod *__restrict__ msg_hash, const uint32_t *__restrict__ msg_idx,
                                                      size_t nattest, size_t nkeys, size_t nmsg, g1_affine_pod *__restrict__ P,
                                                      g2_affine_pod *__restrict__ Q, uint32_t *__restrict__ group_off,
                                                      uint8_t *__restrict__ valid) {
    size_t a = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a == 0) group_off[0] = 0;
    if (a >= nattest) return;
    group_off[a + 1] = (uint32_t)(2 * a + 2);
    // fail closed (the reference panics on these, pairing.go:17-26 / g1pubs/bls.go:287-290 with no keys): an empty committee, a
    // key or message index outside the tables, keys that cancel to the point at infinity, an infinite signature
    const uint32_t lo = key_off[a], hi = key_off[a + 1];
    bool good = hi > lo && msg_idx[a] < nmsg && !sig[a].inf;
    for (uint32_t k = lo; k < hi && good; k++) good = key_idx[k] < nkeys;
    xyzz<FpInl> acc;
    if (good) msm_bucket_sum(acc, registry, key_idx, lo, hi); else xyzz_set_inf(acc);
    fp ox, oy, oz;
    xyzz_to_jac_normalised(ox, oy, oz, acc);
    g1_affine_pod *neg = P + 2 * a + 1, *one = P + 2 * a;
    bool inf = fp_is_zero(oz);
    fp_neg(oy, oy);
    if (inf) { fp_set_zero(ox); fp_set_one(oy); }     // G1AffineZero = (0, 1, true), g1.go:22
    fp_store_u64(neg->x, ox); fp_store_u64(neg->y, oy);
    neg->inf = inf ? 1 : 0;
    for (int i = 0; i < 7; i++) neg->pad[i] = 0;
    const uint32_t gx[12] = {B381_G1_GEN_X_LIMBS}, gy[12] = {B381_G1_GEN_Y_LIMBS};
    fp t;
    fp_load_tab(t, gx); fp_store_u64(one->x, t);
    fp_load_tab(t, gy); fp_store_u64(one->y, t);
    one->inf = 0;
    for (int i = 0; i < 7; i++) one->pad[i] = 0;
    Q[2 * a] = sig[a];
    Q[2 * a + 1] = msg_hash[msg_idx[a] < nmsg ? msg_idx[a] : 0];
    valid[a] = (good && !inf) ? 1 : 0;
}
// ---- attestation-level random-linear-combination check: helpers ---------------------------------------------------------------
// the group key of attestation a as a scalar: message index + 1 (0 = no bucket; an index outside the table is flagged by k_attest_pairs)
__global__ void k_group_keys(const uint32_t *__restrict__ msg_idx, size_t n, size_t nmsg, uint64_t *__restrict__ k) {
    size_t a = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const uint32_t m = msg_idx[a];
    k[4 * a] = m < nmsg ? (uint64_t)m + 1 : 0; k[4 * a + 1] = 0; k[4 * a + 2] = 0; k[4 * a + 3] = 0;
}
// any attestation that k_attest_pairs rejected, or a weight that is zero or wider than `bits`, makes the whole check false
__global__ void k_attest_rlc_valid(const uint8_t *__restrict__ valid, const uint64_t *__restrict__ r, int bits, size_t n, uint32_t *__restrict__ any_bad) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t any = 0, over = 0;
    for (int l = 0; l < 4; l++) {
        uint64_t w = r[4 * i + l];
        any |= w;
        int lo = 64 * l;
        if (bits <= lo) over |= w; else if (bits < lo + 64) over |= w >> (bits - lo);
    }
    if (!valid[i] || any == 0 || over != 0) atomicOr(any_bad, 1u);
}
// pair m of the check: (-(sum of bucket m + 1), H_m); the buckets hold sums of r_a (-pk_a), so the pair is (sum r_a pk_a, H_m).
// A message no attestation refers to gives the point at infinity (its pair contributes the factor 1).
__global__ void __launch_bounds__(64) k_group_pairs(const xyzz<FpInl> *__restrict__ chunks, const uint32_t *__restrict__ chunk_off, uint32_t nmsg,
                                                    const g2_affine_pod *__restrict__ msg_hash, g1_affine_pod *__restrict__ P,
                                                    g2_affine_pod *__restrict__ Q) {
    uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nmsg) return;
    xyzz<FpInl> acc;
    const uint32_t c0 = chunk_off[m + 1];
    if (chunk_off[m + 2] > c0) acc = chunks[c0]; else xyzz_set_inf(acc);
    fp ox, oy, oz;
    xyzz_to_jac_normalised(ox, oy, oz, acc);
    const bool inf = fp_is_zero(oz);
    fp_neg(oy, oy);
    if (inf) { fp_set_zero(ox); fp_set_one(oy); }     // G1AffineZero = (0, 1, true), g1.go:22
    fp_store_u64(P[m].x, ox); fp_store_u64(P[m].y, oy);
    P[m].inf = inf ? 1 : 0;
    for (int i = 0; i < 7; i++) P[m].pad[i] = 0;
    Q[m] = msg_hash[m];
}
// ok[i] &= valid[i]
__global__ void k_and_bytes2(uint8_t *__restrict__ ok, const uint8_t *__restrict__ valid, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ok[i] = (ok[i] && valid[i]) ? 1 : 0;
}

#endif  // __CUDACC__

}  // namespace b381
