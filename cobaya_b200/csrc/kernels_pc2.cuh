// kernels_pc2.cuh -- k_step_pc2: the producer/consumer Metropolis step kernel with the
// proposal-side products SPLIT ACROSS THE FOUR SM SUB-PARTITIONS (sm_100a).
//
// Same algorithm and same consumer as k_step_pc (kernels_fast.cuh; mcmc.py:545-562,
// 670-748, proposal.py:206-224, gaussian_mixture.py:138-163), for one block of all D
// parameters, one Gaussian mode and a likelihood matrix that is lower-triangular in sorted
// coordinates.  What changes is who forms the products.  With
//     delta^ = T u            (proposal.py:224, u = column k of the Haar basis)
//     w^     = G u,  G = L^-1 P T   (lower triangular; formed once on the host)
// both products of a proposal read the SAME direction u and are independent of each other
// and of the chain state; the radius r multiplies them in the consumer
// (x' = x + (r s) delta^, y' = y + (r s) w^, one FMA each: no extra work).
//
// k_step_pc gives each tile of 8 chains its own producer warp: 7 tiles per SM on 4
// sub-partitions leave the FP64 tensor pipe of the SM 2-2-2-1 loaded, and every B fragment
// of T and L^-1 P is re-read from shared memory for every tile.  Here an SM has
//   * 2 x 4 (= 2 (NT+1)/2) producer warps, two per sub-partition.  Producers j and j+4 own
//     the OUTPUT row tiles {NT-1-j, j} of both products -- 4 (NT+1) m8n8k4 DMMAs per
//     proposal tile, an exactly even split of the triangular work over the four
//     sub-partitions -- and take the even / the odd tile-steps of ALL tiles of the SM: while
//     one of them goes through the latencies at an item boundary (mbarrier tests, the
//     direction loads, draining the pipe, stores), the other keeps the tensor pipe of their
//     sub-partition busy.  Their B fragments (2 (NT+1) blocks) never change: they live in
//     REGISTERS for the whole window, so the inner loop has no B-fragment loads at all;
//   * 1 loader warp: reads the plan of the window and brings each tile-step's 8 direction
//     vectors (8 x 8D bytes, contiguous in the basis store) into a shared-memory ring with
//     cp.async (LDGSTS) tracked by an mbarrier, several items ahead; the constant block is
//     staged with one TMA bulk copy (cp.async.bulk);
//   * 1 consumer warp per tile: bounds, prior, |y + w|^2, Metropolis test, bookkeeping, row
//     store -- k_step_pc's consumer with FMAs in place of the adds.
// Rings: directions (DRING items, full: the loader's copies, empty: all producers), products (per
// tile 2 slots, full: all producers, empty: the tile's consumer).
#pragma once
#include "kernels_fast.cuh"

#ifndef CB2_PC2_DRING
#define CB2_PC2_DRING 8     // direction ring: items (tile-steps) in flight
#endif
#define CB2_PC2_ORING 2     // product ring: slots per tile
#define CB2_PC2_MAX_TILES 7   // 1 loader + 8 producers + 7 consumers = 16 warps

__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes,
                                             uint32_t mbar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar)
        : "memory");
}

static inline bool pc2_step_supported(const ModelDev &M, const FastPackDesc &P, int n_blocks,
                                      int D) {
    return pc_step_supported(M, P) && P.tri_like && n_blocks == 1 && (D % 2) == 0 && P.NT >= 2 &&
           P.iofj_identity;
}

// cycle counters of one consumer / one producer warp (CTA 0), filled by -DCB2_PC2_TIMING builds
__device__ long long g_pc2_dbg[16];
#ifdef CB2_PC2_TIMING
#define PC2_T(var) const long long var = clock64()
#define PC2_ACC(i, a, b) do { if (dbg) g_pc2_dbg[i] += (b) - (a); } while (0)
#else
#define PC2_T(var)
#define PC2_ACC(i, a, b)
#endif

template <int NT> struct Pc2Layout {
    static constexpr int DP = NT * 8;
    static constexpr int LDR = DP + ((NT % 2 == 0) ? 8 : 0);  // row stride: 64 B mod 128 B
    static constexpr int DITEM = 8 * LDR;                     // doubles per direction item
    static constexpr int OSLOT = 2 * NT * 64;                 // delta^ + w^, fragment order
    static constexpr int NPW = (NT + 1) / 2;                  // producer warps per phase
};

// ---- producer warp J: output row tiles A = NT-1-J and B = J of both products
template <int NT, int J>
__device__ __forceinline__ void pc2_producer(const double *__restrict__ gT,
                                             const double *__restrict__ gG, int lane,
                                             int phase, int n_items, int ntile,
                                             const double *dring,
                                             double *oring, uint32_t mb_dfull,
                                             uint32_t mb_dempty, uint32_t mb_ofull,
                                             uint32_t mb_oempty, uint32_t mb_burst,
                                             uint32_t mb_stepdone) {
    using Ly = Pc2Layout<NT>;
    constexpr int A = NT - 1 - J, B = J;
    constexpr bool TWO = (A != B);
    const int q = lane >> 2, r = lane & 3;
    // B fragments of this warp's rows, resident in registers for the whole window
    double2 tA[A + 1], gA[A + 1], tB[B + 1], gB[B + 1];
    {
        const double2 *t2 = reinterpret_cast<const double2 *>(gT) + lane;
        const double2 *g2 = reinterpret_cast<const double2 *>(gG) + lane;
#pragma unroll
        for (int m = 0; m <= A; ++m) {
            tA[m] = __ldg(t2 + ((A * (A + 1)) / 2 + m) * 32);
            gA[m] = __ldg(g2 + ((A * (A + 1)) / 2 + m) * 32);
        }
#pragma unroll
        for (int m = 0; m <= B; ++m) {
            tB[m] = __ldg(t2 + ((B * (B + 1)) / 2 + m) * 32);
            gB[m] = __ldg(g2 + ((B * (B + 1)) / 2 + m) * 32);
        }
    }
    // this warp takes the items idx = phase, phase + 2, ... (item = tile-step, tile fastest)
    int t = phase % ntile, s = phase / ntile;
    int s_cleared = 0;  // steps < s_cleared: the consumers' FP64 bursts are over
#ifdef CB2_PC2_TIMING
    const bool dbg = blockIdx.x == 0 && J == 0 && phase == 0 && lane == 0;
#endif
    for (int idx = phase; idx < n_items; idx += 2) {
        PC2_T(p0);
#ifdef CB2_PC2_PHASED
        // No DMMA of step s while the consumers are in the vector-FP64 burst of step s-1: on
        // this pipe a DFMA issued between DMMAs costs about as much as a DMMA (the pipe drains
        // at every switch; tools/fp64_mix.cu), one issued while the DMMAs pause costs 2-3 cycles
        while (s_cleared < s) {
            mbar_wait(mb_burst, (uint32_t)s_cleared & 1u);
            ++s_cleared;
        }
#endif
        const int dslot = idx % CB2_PC2_DRING;
        const uint32_t duse = (uint32_t)(idx / CB2_PC2_DRING);
#if !defined(CB2_PC2_NO_LOAD) && !defined(CB2_PC2_NO_DWAIT)
        mbar_wait(mb_dfull + 8u * dslot, duse & 1u);
#endif
        PC2_T(p1);
        const double *urow = dring + (size_t)dslot * Ly::DITEM + q * Ly::LDR + 2 * r;
        // four independent DMMA chains (two products x two row tiles); the partner warp of
        // this sub-partition covers what dependent-issue latency is left
        double dA0 = 0, dA1 = 0, wA0 = 0, wA1 = 0, dB0 = 0, dB1 = 0, wB0 = 0, wB1 = 0;
#ifdef CB2_PC2_NO_MMA
        constexpr int MTOP = -1;
#else
        constexpr int MTOP = A;
#endif
#pragma unroll
        for (int m = 0; m <= MTOP; ++m) {
            const double2 u = *reinterpret_cast<const double2 *>(urow + 8 * m);
            dmma8x8x4(dA0, dA1, u.x, tA[m].x);
            dmma8x8x4(wA0, wA1, u.x, gA[m].x);
            if (TWO && m <= B) {
                dmma8x8x4(dB0, dB1, u.x, tB[m].x);
                dmma8x8x4(wB0, wB1, u.x, gB[m].x);
            }
            dmma8x8x4(dA0, dA1, u.y, tA[m].y);
            dmma8x8x4(wA0, wA1, u.y, gA[m].y);
            if (TWO && m <= B) {
                dmma8x8x4(dB0, dB1, u.y, tB[m].y);
                dmma8x8x4(wB0, wB1, u.y, gB[m].y);
            }
        }
        // the direction slot is free as soon as this warp's loads are done
#ifndef CB2_PC2_NO_LOAD
        __syncwarp();
        if (lane == 0) mbar_arrive(mb_dempty + 8u * dslot);
#endif
        PC2_T(p2);
        const int oslot = s % CB2_PC2_ORING;
        const uint32_t ouse = (uint32_t)(s / CB2_PC2_ORING);
        const uint32_t ob = (uint32_t)(t * CB2_PC2_ORING + oslot);
        if (ouse > 0) mbar_wait(mb_oempty + 8u * ob, (ouse - 1) & 1u);
        PC2_T(p3);
        double2 *sl = reinterpret_cast<double2 *>(oring + (size_t)ob * Ly::OSLOT) + lane;
        sl[A * 32] = make_double2(dA0, dA1);
        sl[(NT + A) * 32] = make_double2(wA0, wA1);
        if (TWO) {
            sl[B * 32] = make_double2(dB0, dB1);
            sl[(NT + B) * 32] = make_double2(wB0, wB1);
        }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(mb_ofull + 8u * ob);
#ifdef CB2_PC2_PHASED
            mbar_arrive(mb_stepdone);
#endif
        }
        t += 2;
        while (t >= ntile) { t -= ntile; ++s; }
        PC2_T(p4);
        PC2_ACC(8, p0, p1);   // burst + dfull waits
        PC2_ACC(9, p1, p2);   // LDS + DMMA issue
        PC2_ACC(10, p2, p3);  // oempty wait
        PC2_ACC(11, p3, p4);  // stores, arrive
    }
}

template <int NT, bool HAS_NORMAL>
__global__ void __launch_bounds__((1 + 2 * Pc2Layout<NT>::NPW + CB2_PC2_MAX_TILES) * 32, 1)
k_step_pc2(ModelDev M, ChainState S, WindowDev W, const double *__restrict__ gpack,
           const double *__restrict__ gG, FastPackDesc P, const double2 *__restrict__ draws,
           const int2 *__restrict__ plan, int64_t n_chains, int n_steps, int ntile_max,
           int64_t n_tiles) {
    using Ly = Pc2Layout<NT>;
    constexpr int NPW = Ly::NPW;
    extern __shared__ __align__(16) double fsm[];
    __shared__ __align__(8) unsigned long long mbar_pack, mbar_burst, mbar_stepdone;
    __shared__ __align__(8) unsigned long long mbar_dfull[CB2_PC2_DRING], mbar_dempty[CB2_PC2_DRING];
    __shared__ __align__(8) unsigned long long mbar_ofull[CB2_PC2_MAX_TILES * CB2_PC2_ORING],
        mbar_oempty[CB2_PC2_MAX_TILES * CB2_PC2_ORING];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // tiles of this CTA: an even split of n_tiles over the grid (ntile <= ntile_max)
    const int64_t tile_lo = (n_tiles * blockIdx.x) / gridDim.x,
                  tile_hi = (n_tiles * (blockIdx.x + 1)) / gridDim.x;
    const int ntile = (int)(tile_hi - tile_lo);
    if (ntile <= 0) return;
    // shared memory: [constants from off_A on][direction ring][product ring]
    const int pack_lo = P.off_A;
    const int pack_n = P.total - pack_lo;
    double *pack = fsm - pack_lo;  // pack[P.off_x + ...] addresses as in k_step_pc
    double *dring = fsm + ((pack_n + 1) & ~1);
    double *oring = dring + CB2_PC2_DRING * Ly::DITEM;
    const uint32_t mb_pack = (uint32_t)__cvta_generic_to_shared(&mbar_pack);
    const uint32_t mb_burst = (uint32_t)__cvta_generic_to_shared(&mbar_burst);
    const uint32_t mb_stepdone = (uint32_t)__cvta_generic_to_shared(&mbar_stepdone);
    const uint32_t mb_dfull = (uint32_t)__cvta_generic_to_shared(mbar_dfull);
    const uint32_t mb_dempty = (uint32_t)__cvta_generic_to_shared(mbar_dempty);
    const uint32_t mb_ofull = (uint32_t)__cvta_generic_to_shared(mbar_ofull);
    const uint32_t mb_oempty = (uint32_t)__cvta_generic_to_shared(mbar_oempty);
    if (tid == 0) {
        mbar_init(mb_pack, 1);
        mbar_init(mb_burst, (uint32_t)ntile);  // one arrival per consumer warp and step
        mbar_init(mb_stepdone, (uint32_t)(NPW * ntile));  // NPW producer warps per tile-step
        for (int i = 0; i < CB2_PC2_DRING; ++i) {
            mbar_init(mb_dfull + 8u * i, 32);  // one noinc arrival per loader lane
            mbar_init(mb_dempty + 8u * i, NPW);
        }
        for (int i = 0; i < CB2_PC2_MAX_TILES * CB2_PC2_ORING; ++i) {
            mbar_init(mb_ofull + 8u * i, NPW);
            mbar_init(mb_oempty + 8u * i, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // padding columns of the direction rows stay zero for the whole kernel
    for (int e = tid; e < CB2_PC2_DRING * Ly::DITEM; e += blockDim.x) dring[e] = 0.0;
    // order the generic-proxy zero fill before the async-proxy (TMA) writes that follow
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(mb_pack, (uint32_t)pack_n * 8u);
        tma_bulk_g2s((uint32_t)__cvta_generic_to_shared(fsm), gpack + pack_lo,
                     (uint32_t)pack_n * 8u, mb_pack);
    }
    const int n_items = n_steps * ntile;
    const int D = M.D;

    if (warp < 2 * NPW) {
        // ================================ producers ===================================
        // warps j and j + NPW share the row tiles j (and, for NPW = 4, the sub-partition)
        const double *gT = gpack + P.off_T;
        const int phase = warp / NPW;
#define CB2_PC2_P(JJ)                                                                       \
    case JJ:                                                                                \
        if constexpr (JJ < NPW)                                                             \
            pc2_producer<NT, JJ>(gT, gG, lane, phase, n_items, ntile, dring, oring,         \
                                 mb_dfull, mb_dempty, mb_ofull, mb_oempty, mb_burst,        \
                                 mb_stepdone);                                              \
        break;
        switch (warp % NPW) { CB2_PC2_P(0) CB2_PC2_P(1) CB2_PC2_P(2) CB2_PC2_P(3) }
#undef CB2_PC2_P
    } else if (warp == 2 * NPW) {
        // ================================ loader ======================================
        // One tile-step = 8 direction vectors of D doubles, each contiguous in the basis store
        // (row pl.x of the chain's bases).  They are brought in with cp.async (LDGSTS, 16 B
        // per lane: one instruction per chain) and tracked by the slot's mbarrier
        // (cp.async.mbarrier.arrive.noinc: the barrier completes when the copies of all 32
        // lanes have landed).  Eight 512-byte TMA bulk copies per item were tried first: the
        // kernel then ran at the rate the copy engine accepts small copies (~1 item per
        // 970 cycles, tensor pipe 61 % busy even without consumers).
        const int nb = D;
        const double *basis = W.basis[0];
        const uint32_t dring_a = (uint32_t)__cvta_generic_to_shared(dring);
        constexpr int PF = 4;  // plan entries are fetched PF items ahead
        int2 pq[PF];
        auto plan_of = [&](int idx) -> int2 {
            const int s = idx / ntile, t = idx - s * ntile;
            int64_t chain = (tile_lo + t) * 8 + (lane & 7);
            if (chain >= n_chains) chain = n_chains - 1;
            return ldg_int2_early(plan + chain * (int64_t)n_steps + s);
        };
#pragma unroll
        for (int k = 0; k < PF; ++k) pq[k] = (k < n_items) ? plan_of(k) : make_int2(0, 0);
#ifdef CB2_PC2_NO_LOAD
        if (n_items > 0) return;
#endif
        for (int idx0 = 0; idx0 < n_items; idx0 += PF) {
#pragma unroll
            for (int k = 0; k < PF; ++k) {
                const int idx = idx0 + k;
                if (idx < n_items) {
                    const int dslot = idx % CB2_PC2_DRING;
                    const uint32_t duse = (uint32_t)(idx / CB2_PC2_DRING);
                    if (duse > 0) mbar_wait(mb_dempty + 8u * dslot, (duse - 1) & 1u);
                    const int2 pl = pq[k];
                    if (idx + PF < n_items) pq[k] = plan_of(idx + PF);
                    const uint32_t dst0 = dring_a + (uint32_t)((dslot * Ly::DITEM + 2 * lane) * 8);
#pragma unroll
                    for (int qq = 0; qq < 8; ++qq) {
                        const int row = __shfl_sync(0xffffffffu, pl.x, qq);
                        const double *src = basis + (size_t)row * (size_t)nb + 2 * lane;
                        if (2 * lane < nb)
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                                             dst0 + (uint32_t)(qq * Ly::LDR * 8)),
                                         "l"(src)
                                         : "memory");
                    }
                    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(
                                     mb_dfull + 8u * dslot)
                                 : "memory");
                }
            }
        }
    } else {
        // ================================ consumers ===================================
        const int t = warp - 2 * NPW - 1;
        if (t >= ntile) return;
        mbar_wait(mb_pack, 0);
        const int q = lane >> 2, r = lane & 3;
        const int64_t chain_raw = (tile_lo + t) * 8 + q;
        const bool active = chain_raw < n_chains;
        const int64_t chain = active ? chain_raw : (n_chains - 1);
        const double2 *my_draws = draws + chain * (int64_t)n_steps;
        const long long *klo = reinterpret_cast<const long long *>(pack + P.off_klo),
                        *kup = reinterpret_cast<const long long *>(pack + P.off_kup);
        const int *pflag = reinterpret_cast<const int *>(pack + P.off_flags);
        double xs[NT][2], ys[NT][2];
        uint32_t m_norm = 0;
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            const int j = 8 * n + 2 * r;
            // sorted == sampler order (iofj_identity); D even: pairs never straddle D
            if (j < D) {
                const double2 v2 = *reinterpret_cast<const double2 *>(S.x + chain * D + j);
                xs[n][0] = v2.x; xs[n][1] = v2.y;
            } else {
                xs[n][0] = 0.0; xs[n][1] = 0.0;
            }
            m_norm |= (uint32_t)(pflag[j] & 1) << (2 * n);
            m_norm |= (uint32_t)(pflag[j + 1] & 1) << (2 * n + 1);
        }
        {   // y = L^-1 P (x - mu), refreshed at every window start
            const double *mu = pack + P.off_mu;
            double z[NT][2];
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                const double2 m2 = *reinterpret_cast<const double2 *>(mu + 8 * n + 2 * r);
                z[n][0] = xs[n][0] - m2.x;
                z[n][1] = xs[n][1] - m2.y;
                ys[n][0] = 0.0;
                ys[n][1] = 0.0;
            }
            warp_matvec8<NT, true>(pack + P.off_A, lane, z, ys);
        }
        double logpost = S.logpost[chain], logprior = S.logprior[chain], loglike = S.ll[chain];
        long long weight = S.weight[chain], prior_rej = S.prior_rej[chain],
                  burn_left = S.burn_left[chain], added_w = S.added_w[chain],
                  n_rows = S.n_rows[chain], n_acc = S.n_acc[chain];
        uint32_t flags = S.flags[chain];
        const double c0 = pack[P.off_c0];
        const double scale = M.proposal_scale;
        double2 dr_next = ldg_f64x2_early(my_draws);
#ifdef CB2_PC2_TIMING
        const bool dbg = blockIdx.x == 0 && t == 0 && lane == 0;
#endif
        for (int s = 0; s < n_steps; ++s) {
            PC2_T(c0t);
            const int oslot = s % CB2_PC2_ORING;
            const uint32_t ouse = (uint32_t)(s / CB2_PC2_ORING);
            const uint32_t ob = (uint32_t)(t * CB2_PC2_ORING + oslot);
            const double rs = dr_next.x * scale, e_acc = dr_next.y;
            if (s + 1 < n_steps) dr_next = ldg_f64x2_early(my_draws + s + 1);
            mbar_wait(mb_ofull + 8u * ob, ouse & 1u);
#ifdef CB2_PC2_PHASED
            // every tile of this step has been produced: from here until the consumers'
            // arrivals on mb_burst no DMMA is in flight on this SM, and the vector-FP64
            // instructions below issue at their own rate
            mbar_wait(mb_stepdone, (uint32_t)s & 1u);
#endif
            PC2_T(c1t);
#ifdef CB2_PC2_NO_CONSUME
            if (n_steps > 0) {
                __syncwarp();
                if (lane == 0) { mbar_arrive(mb_burst); mbar_arrive(mb_oempty + 8u * ob); }
                continue;
            }
#endif
            double2 *sl = reinterpret_cast<double2 *>(oring + (size_t)ob * Ly::OSLOT);
            // One warp carries a tile, so the step is bound by LATENCY: no long dependent
            // chains.  |y'|^2 goes to four partial sums; the bounds are tested on the integer
            // pipe with order-preserving keys (key(lo) <= key(x) <= key(up), NaN and the
            // infinities fall outside or are caught by the exponent test), as independent
            // flag words instead of one chain of FP64 set-predicate instructions.
            unsigned badm = 0;
            double ps = 0.0, q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                const double2 d2 = sl[n * 32 + lane];
                const double2 w2 = sl[(NT + n) * 32 + lane];
                const longlong2 lo2 = *reinterpret_cast<const longlong2 *>(klo + 8 * n + 2 * r);
                const longlong2 up2 = *reinterpret_cast<const longlong2 *>(kup + 8 * n + 2 * r);
                const double x0 = fma(rs, d2.x, xs[n][0]), x1 = fma(rs, d2.y, xs[n][1]);
                const double y0 = fma(rs, w2.x, ys[n][0]), y1 = fma(rs, w2.y, ys[n][1]);
                if (n & 1) { q2 = fma(y0, y0, q2); q3 = fma(y1, y1, q3); }
                else { q0 = fma(y0, y0, q0); q1 = fma(y1, y1, q1); }
                // the candidate point replaces the products in the slot: an accepted step
                // takes it from there without touching the FP64 pipe again
                sl[n * 32 + lane] = make_double2(x0, x1);
                sl[(NT + n) * 32 + lane] = make_double2(y0, y1);
                long long k0 = __double_as_longlong(x0), k1 = __double_as_longlong(x1);
                const unsigned e0 = (unsigned)(k0 >> 52) & 0x7ffu, e1 = (unsigned)(k1 >> 52) & 0x7ffu;
                k0 ^= (k0 >> 63) & 0x7fffffffffffffffLL;
                k1 ^= (k1 >> 63) & 0x7fffffffffffffffLL;
                badm |= (unsigned)(k0 < lo2.x) | (unsigned)(k0 > up2.x) | (unsigned)(e0 == 0x7ffu) |
                        (unsigned)(k1 < lo2.y) | (unsigned)(k1 > up2.y) | (unsigned)(e1 == 0x7ffu);
                if (HAS_NORMAL) {
                    if ((m_norm >> (2 * n)) & 1u) {
                        const int j = 8 * n + 2 * r;
                        const double zz = (x0 - pack[P.off_loc + j]) / pack[P.off_isc + j];
                        ps += pack[P.off_mls + j] - zz * zz / 2;
                    }
                    if ((m_norm >> (2 * n + 1)) & 1u) {
                        const int j = 8 * n + 2 * r + 1;
                        const double zz = (x1 - pack[P.off_loc + j]) / pack[P.off_isc + j];
                        ps += pack[P.off_mls + j] - zz * zz / 2;
                    }
                }
            }
            PC2_T(c2t);
#ifdef CB2_PC2_PHASED
            // the bulk of this step's vector-FP64 work is done: the producers may go on
            __syncwarp();
            if (lane == 0) mbar_arrive(mb_burst);
#endif
            badm |= __shfl_xor_sync(0xffffffffu, badm, 1);
            badm |= __shfl_xor_sync(0xffffffffu, badm, 2);
            const bool bad = badm != 0;
            double qsum = quad_sum((q0 + q1) + (q2 + q3));
            if (HAS_NORMAL) ps = quad_sum(ps);
            const double t_prior = bad ? -CUDART_INF : (M.uniform_logp + ps);
            const double t_like = -0.5 * (c0 + qsum);
            const double t_post = bad ? -CUDART_INF : (t_prior + t_like);
            bool acc;
            if (t_post == -CUDART_INF) acc = false;
            else if (t_post > logpost) acc = true;
            else acc = (M.temperature == 1.0) ? (e_acc > logpost - t_post)
                                              : (e_acc > (logpost - t_post) / M.temperature);
            PC2_T(c3t);
            if (acc) {
                if (burn_left <= 0) {
                    long long wst = weight;
                    bool store = true;
                    if (M.output_thin > 1) {
                        added_w += weight;
                        if (added_w >= M.output_thin) {
                            wst = added_w / M.output_thin;
                            added_w %= M.output_thin;
                        } else store = false;
                    }
                    if (store) {
                        if (n_rows >= S.cap) flags |= CB2_FLAG_ROWS_FULL;
                        else {
                            if (active) {
                                double *row = S.rows + ((size_t)chain * S.cap + n_rows) * M.width;
                                if (r == 0) {
                                    row[0] = (double)wst;
                                    row[1] = (M.temperature == 1.0) ? -logpost
                                                                    : -(logpost / M.temperature);
                                } else if (r == 1) {
                                    row[2 + D] = -logprior;
                                    row[3 + D] = -logprior;
                                } else if (r == 2) {
                                    row[4 + D] = -2 * loglike;
                                    row[5 + D] = -2 * loglike;
                                }
#pragma unroll
                                for (int n = 0; n < NT; ++n)
                                    if (8 * n + 2 * r < D)
                                        *reinterpret_cast<double2 *>(row + 2 + 8 * n + 2 * r) =
                                            make_double2(xs[n][0], xs[n][1]);
                            }
                            n_rows += 1;
                        }
                    }
                } else burn_left -= 1;
#pragma unroll
                for (int n = 0; n < NT; ++n) {  // the candidate evaluated above
                    const double2 xn = sl[n * 32 + lane];
                    const double2 yn = sl[(NT + n) * 32 + lane];
                    xs[n][0] = xn.x; xs[n][1] = xn.y;
                    ys[n][0] = yn.x; ys[n][1] = yn.y;
                }
                logpost = t_post; logprior = t_prior; loglike = t_like;
                weight = 1; prior_rej = 0; n_acc += 1;
            } else {
                weight += 1;
                if (t_prior == -CUDART_INF) prior_rej += 1;
                const long long sgn = (burn_left > 0) - (burn_left < 0);
                if (weight - prior_rej > M.max_tries * (1 + 9 * sgn)) flags |= CB2_FLAG_STUCK;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(mb_oempty + 8u * ob);
            PC2_T(c4t);
            PC2_ACC(0, c0t, c1t);  // wait for the products
            PC2_ACC(1, c1t, c2t);  // bulk: loads, FMAs, bounds
            PC2_ACC(2, c2t, c3t);  // reductions, accept test
            PC2_ACC(3, c3t, c4t);  // bookkeeping, store, update, release
        }
        __syncwarp();
        if (active) {
#pragma unroll
            for (int n = 0; n < NT; ++n)
                if (8 * n + 2 * r < D)
                    *reinterpret_cast<double2 *>(S.x + chain * D + 8 * n + 2 * r) =
                        make_double2(xs[n][0], xs[n][1]);
            if (r == 0) {
                S.logpost[chain] = logpost; S.logprior[chain] = logprior; S.ll[chain] = loglike;
                S.weight[chain] = weight; S.prior_rej[chain] = prior_rej;
                S.burn_left[chain] = burn_left; S.added_w[chain] = added_w;
                S.n_rows[chain] = n_rows; S.n_acc[chain] = n_acc;
                if (flags) atomicOr(&S.flags[chain], flags);
            }
        }
    }
}

template <int NT>
static int launch_step_pc2_t(cudaStream_t st, const ModelDev &M, const ChainState &S,
                             const WindowDev &W, const double *gpack, const double *gG,
                             const FastPackDesc &P, const double2 *draws, const int2 *plan,
                             int64_t n_chains, int n_steps, int sm_count) {
    using Ly = Pc2Layout<NT>;
    const int64_t tiles = (n_chains + 7) / 8;
    int wpc = (int)((tiles + sm_count - 1) / sm_count);
    if (wpc < 1) wpc = 1;
    if (wpc > CB2_PC2_MAX_TILES) wpc = CB2_PC2_MAX_TILES;
    const int grid = (int)((tiles + wpc - 1) / wpc);
    const int pack_n = P.total - P.off_A;
    const size_t smem = ((size_t)((pack_n + 1) & ~1) + (size_t)CB2_PC2_DRING * Ly::DITEM +
                         (size_t)wpc * CB2_PC2_ORING * Ly::OSLOT) * 8;
    if (smem > 226 * 1024) return -2;
    const int threads = (1 + 2 * Ly::NPW + wpc) * 32;
    cudaError_t e;
    if (M.any_normal) {
        e = cudaFuncSetAttribute(k_step_pc2<NT, true>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return -1000 - (int)e;
        k_step_pc2<NT, true><<<grid, threads, smem, st>>>(M, S, W, gpack, gG, P, draws, plan,
                                                          n_chains, n_steps, wpc, tiles);
    } else {
        e = cudaFuncSetAttribute(k_step_pc2<NT, false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return -1000 - (int)e;
        k_step_pc2<NT, false><<<grid, threads, smem, st>>>(M, S, W, gpack, gG, P, draws, plan,
                                                           n_chains, n_steps, wpc, tiles);
    }
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -2000 - (int)e;
}

static inline int launch_step_pc2(cudaStream_t st, const ModelDev &M, const ChainState &S,
                                  const WindowDev &W, const double *gpack, const double *gG,
                                  const FastPackDesc &P, const double2 *draws, const int2 *plan,
                                  int64_t n_chains, int n_steps, int sm_count) {
#define CB2_PC2(N)                                                                          \
    case N:                                                                                 \
        return launch_step_pc2_t<N>(st, M, S, W, gpack, gG, P, draws, plan, n_chains,       \
                                    n_steps, sm_count);
    switch (P.NT) { CB2_PC2(2) CB2_PC2(3) CB2_PC2(4) CB2_PC2(5) CB2_PC2(6) CB2_PC2(7) CB2_PC2(8) }
#undef CB2_PC2
    return -2;
}
