// policy_tc_kernel.cu — CommCategoricalMLPPolicy forward on the 5th-generation tensor cores (tcgen05).
//
// Same formula and outputs as policy_kernel.cu (see its header for the reference lines), different machine
// mapping: a CTA of 512 threads owns a tile of 128 agent rows (whole environments, n <= 64) and runs
//   * every row-wise dense layer (encoder, attention query, H_l Wg_l, the categorical head) as
//     tcgen05.mma.kind::f16 (fp16 operands, fp32 accumulators in tensor memory).  fp32-level accuracy is kept by
//     error compensation: x = x_hi + 2^-12 x_lo with x_hi = fp16(x), x_lo = fp16((x - x_hi) * 4096);
//     A_hi x [B_hi ; B_lo] (B stacked along N) gives hi*hi and hi*lo in two accumulators with one series of K/16
//     instructions, A_lo x B_hi adds the other cross term, the epilogue forms acc0 + 2^-12 acc1 (dropped term:
//     2^-24).  Measured error of one product 1-3e-7 of scale, i.e. fp32 level.  Weights are pre-split, stacked and
//     pre-arranged in the canonical K-major core-matrix layout by cm_policy_tc_prepare(), so that a layer's
//     B operand is ONE bulk async copy (TMA engine, mbarrier completion) issued one to two layers ahead;
//   * the per-environment pieces (n x n scores, softmax, masked renormalisation, aggregation over
//     neighbours) exactly — not as padded tile products — on the CUDA cores from k-major fp32 copies of E, Q
//     and H_l Wg_l in shared memory.
// TMEM lane = tile row = thread (row = 32 * (warp % 4) + lane); the four warps that share a lane quadrant
// split the accumulator columns.  Epilogues read TMEM with tcgen05.ld, apply bias + tanh, and write the next
// A operand (hi / lo) straight into the canonical layout.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "commarl_b200.h"
#include "common.cuh"
#include "policy_layout.cuh"
#include "tc_common.cuh"

#ifndef CM_TC_OBS_PREFETCH
#define CM_TC_OBS_PREFETCH 0   // measured slower on B200 (C2 0.082 vs 0.076 ms): the held registers cost more than the hidden latency
#endif
#ifndef CM_TC_DEBUG
#define CM_TC_DEBUG 0   // timing experiments only: 1 skip MMAs, 2 skip tanh, 4 skip attention loops, 8 skip weight copies
#endif

namespace cm {

using namespace tc;

static constexpr int kTcRows = 128;
static constexpr int kTcThreads = 512;
static constexpr int kActBytes = 32768, kWBytes = 65536;   // A operand hi|lo (128 x 64 fp16 each); weight ring 2 x 32 KB
static constexpr int kTPitch = 128;     // floats per k-major row of ET / QT / HWT

struct TcArgs {
    cm_policy_desc d;
    cm_policy_io io;
    TcPlan plan;
    int envs_per_tile;
    int64_t n_tiles;
};

template <int CW>
__device__ __forceinline__ void ld_cols(uint32_t taddr, float (&v)[CW])
{
    static_assert(CW == 8 || CW == 16, "column chunk");
#pragma unroll
    for (int c = 0; c < CW; c += 8) {
        float t[8];
        tmem_ld8(taddr + (uint32_t)c, t);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[c + i] = t[i];
    }
    tmem_ld_wait();
}

// fp16 split of one value: hi = fp16(x), lo = fp16((x - hi) * 4096)
__device__ __forceinline__ void split16(float x, __half &hi, __half &lo)
{
    hi = __float2half_rn(x);
    lo = __float2half_rn((x - __half2float(hi)) * 4096.0f);
}

// write CW (8 or 16) consecutive columns, starting at local column c0 (a multiple of 8) of a Kp-wide panel, of row `row`
// as the next A operand: hi block, then the lo block 128 * Kp halves further
template <int CW>
__device__ __forceinline__ void write_act(unsigned char *act, int Kp, int row, int c0, const float (&v)[CW])
{
    const uint32_t lo_off = (uint32_t)kTcRows * Kp * 2;
#pragma unroll
    for (int g = 0; g < CW; g += 8) {
        __align__(16) __half h[8], l[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) split16(v[g + i], h[i], l[i]);
        const uint32_t off = canon_off16(row, c0 + g, Kp);
        *reinterpret_cast<uint4 *>(act + off) = *reinterpret_cast<const uint4 *>(h);
        *reinterpret_cast<uint4 *>(act + lo_off + off) = *reinterpret_cast<const uint4 *>(l);
    }
}

// tanh through ex2.approx / rcp.approx: |error| <= ~3e-7 absolute (two units of fp32 rounding at 1.0), an order
// of magnitude below what the 3xTF32 products contribute and ~4x fewer instructions than tanhf.
__device__ __forceinline__ float tanh_fast(float x)
{
    if (CM_TC_DEBUG & 2) return x;
    const float e = __expf(2.0f * x);
    return 1.0f - __fdividef(2.0f, e + 1.0f);
}

// one converged warp: D[:, 0:N] (+)= A_hi B_hi^T, D[:, N:2N] (+)= A_hi B_lo^T + A_lo B_hi^T.  All lanes run the loop
// (descriptors stay warp-uniform), the elected lane issues.
__device__ __forceinline__ void issue_layer(uint32_t d_tmem, const unsigned char *act, const unsigned char *wblk, int N, int Kp,
                                            uint32_t accumulate)
{
    const uint32_t a_hi = smem_u32(act), a_lo = a_hi + (uint32_t)kTcRows * Kp * 2, b = smem_u32(wblk);
    const int nk = Kp >> 4;
    {   // A_hi x [B_hi ; B_lo]  (N' = 2N)
        const uint32_t idesc = make_idesc_f16(kTcRows, 2 * N);
        uint64_t da = make_smem_desc16(a_hi, Kp, 0), db = make_smem_desc16(b, Kp, 0);
        uint32_t acc = accumulate;
#pragma unroll 2
        for (int j = 0; j < nk; ++j) {
            if (!(CM_TC_DEBUG & 1) && elect_one()) mma_f16(d_tmem, da, db, idesc, acc);
            acc = 1;
            da += 16;     // next K = 16 slice: start address + 256 bytes (>> 4)
            db += 16;
        }
    }
    {   // A_lo x B_hi  (N' = N) into the cross-term accumulator
        const uint32_t idesc = make_idesc_f16(kTcRows, N);
        uint64_t da = make_smem_desc16(a_lo, Kp, 0), db = make_smem_desc16(b, Kp, 0);
#pragma unroll 2
        for (int j = 0; j < nk; ++j) {
            if (!(CM_TC_DEBUG & 1) && elect_one()) mma_f16(d_tmem + (uint32_t)N, da, db, idesc, 1u);
            da += 16;
            db += 16;
        }
    }
    __syncwarp();
}

// accumulator read-out: acc0 + 2^-12 acc1 for CW columns starting at column `col` of a product whose D block starts at
// `base` and is 2N columns wide
template <int CW>
__device__ __forceinline__ void ld_acc(uint32_t lane_addr, uint32_t base, int N, int col, float (&v)[CW])
{
    float w[CW];
#pragma unroll
    for (int c = 0; c < CW; c += 8) {
        float t0[8], t1[8];
        tmem_ld8(lane_addr + base + (uint32_t)(col + c), t0);
        tmem_ld8(lane_addr + base + (uint32_t)(N + col + c), t1);
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[c + i] = t0[i]; w[c + i] = t1[i]; }
    }
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < CW; ++c) v[c] = fmaf(w[c], 1.0f / 4096.0f, v[c]);
}

// Exact per-environment attention row: thread (row, sub) owns the keys jj = sub, sub + 4, ... of its env (at most
// KT of them), scores = Q[row] . E[key], softmax across the four threads of the row through `red`, result into
// M^T[jj][row] (which overwrites Q^T once every thread has read it).
template <int KT>
__device__ __forceinline__ void scores_softmax(float *QT, const float *ET, float *red, int row, int j0, int sub, int n, bool valid)
{
    float sc[KT];
    const int nk = valid ? (n - sub + 3) / 4 : 0;
#pragma unroll
    for (int t = 0; t < KT; ++t) sc[t] = 0.0f;
    if (nk > 0 && !(CM_TC_DEBUG & 4)) {
#pragma unroll 4
        for (int k = 0; k < 64; ++k) {
            const float qv = QT[k * kTPitch + row];
            const float *er = ET + k * kTPitch + j0 + sub;
#pragma unroll
            for (int t = 0; t < KT; ++t)
                if (t < nk) sc[t] = fmaf(qv, er[4 * t], sc[t]);
        }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < KT; ++t)
        if (t < nk) mx = fmaxf(mx, sc[t]);
    red[sub * kTPitch + row] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[row], red[kTPitch + row]), fmaxf(red[2 * kTPitch + row], red[3 * kTPitch + row]));
    float sum = 0.0f;
#pragma unroll
    for (int t = 0; t < KT; ++t)
        if (t < nk) { sc[t] = expf(sc[t] - mx); sum += sc[t]; }
    red[(4 + sub) * kTPitch + row] = sum;
    __syncthreads();                      // everybody has read QT: it becomes M^T [n][128]
    const float z = red[4 * kTPitch + row] + red[5 * kTPitch + row] + red[6 * kTPitch + row] + red[7 * kTPitch + row];
#pragma unroll
    for (int t = 0; t < KT; ++t)
        if (t < nk) QT[(sub + 4 * t) * kTPitch + row] = sc[t] / z;
    __syncthreads();
}

// bias offsets inside the shared-memory bias table
static constexpr int kBEnc1 = 0, kBEnc2 = 128, kBGcn = 192, kBH1 = 448, kBH2 = 576, kBH3 = 640, kBH4 = 672, kBiasFloats = 704;

struct MmaOp { uint32_t dcol, acc; };

__global__ void __launch_bounds__(kTcThreads, 1) policy_tc_kernel(const TcArgs A)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *ACT = smem;
    unsigned char *WB = smem + kActBytes;                                // weight ring: 2 slots of 32 KB
    float *ET = reinterpret_cast<float *>(smem + kActBytes + kWBytes);   // [64][128] E^T   (keys, residual)
    float *QT = ET + 64 * kTPitch;                                       // [64][128] Q^T, then M^T [n][128]
    float *HWT = QT + 64 * kTPitch;                                      // [64][128] (H_l Wg_l)^T; softmax scratch before that
    float *bias_s = HWT + 64 * kTPitch;                                  // [704]
    uint64_t *bars = reinterpret_cast<uint64_t *>(bias_s + kBiasFloats); // [0],[1] weight slot full, [2] MMAs done
    uint32_t *tmem_s = reinterpret_cast<uint32_t *>(bars + 3);

    const cm_policy_desc &d = A.d;
    const cm_policy_io &io = A.io;
    const TcPlan &P = A.plan;
    const int n = d.n_agents, D = d.obs_dim, L = d.n_layers, W = (n + 31) >> 5;
    const Blob o = blob_layout(D, L);
    const float *__restrict__ wts = io.weights;
    const __half *tcw = reinterpret_cast<const __half *>(io.tc_weights);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int quad = warp & 3, sub = warp >> 2, row = quad * 32 + lane;

    if (warp == 0) tmem_alloc(tmem_s, 512);
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], 1);
        fence_mbar_init();
    }
    for (int i = tid; i < kBiasFloats; i += kTcThreads) {
        float v = 0.0f;
        if (i < kBEnc2) v = wts[o.enc_b1 + i];
        else if (i < kBGcn) v = wts[o.enc_b2 + i - kBEnc2];
        else if (i < kBH1) { if (i - kBGcn < L * 64) v = wts[o.gcn_b + i - kBGcn]; }
        else if (i < kBH2) v = wts[o.head_b1 + i - kBH1];
        else if (i < kBH3) v = wts[o.head_b2 + i - kBH2];
        else if (i < kBH4) v = wts[o.head_b3 + i - kBH3];
        else if (i < kBH4 + CM_ACTIONS) v = wts[o.head_b4 + i - kBH4];
        bias_s[i] = v;
    }
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    const uint32_t tmem = *tmem_s;
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
    const uint32_t DA = 0, DB = 256;          // accumulator column regions (a product's D block is 2N columns wide)
    uint32_t m_phase = 0;
    bool ok = true;

    // ---- weight stream: the stages of the plan are consumed in order, tile after tile; thread 0 keeps the
    // two ring slots full (a slot is refilled as soon as the product that read it has completed) ----
    const uint32_t my_tiles = (uint32_t)((A.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const uint32_t total_blocks = my_tiles * (uint32_t)P.seq_len;
    uint32_t consumed = 0;                    // blocks consumed so far (uniform)
    uint32_t issued = 0, issued_si = 0;       // warp 0: blocks requested so far, and their stage index
    auto issue_loads = [&]() {          // warp 0, converged
        if (warp == 0) {
            while (issued < consumed + 2 && issued < total_blocks) {
                const TcStage &st = P.st[issued_si];
                const uint32_t bytes = (uint32_t)(2 * st.N * st.Kp * 2), slot = issued & 1u;
                if (elect_one()) {
                    if (CM_TC_DEBUG & 8) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[slot])) : "memory"); }
                    else {
                        mbar_expect_tx(&bars[slot], bytes);
                        bulk_g2s(WB + slot * 32768u, tcw + st.w_off, bytes, &bars[slot]);
                    }
                }
                __syncwarp();
                ++issued;
                issued_si = (issued_si + 1 == (uint32_t)P.seq_len) ? 0u : issued_si + 1;
            }
        }
    };
    issue_loads();
    int si = 0;                               // stage index inside the current tile (uniform)
    // all threads: ACT is written -> thread 0 issues one product per op, each against the next weight block ->
    // everybody waits for their completion -> the freed ring slots are refilled
    auto run_mma = [&](int nops, MmaOp op0, MmaOp op1) {
#ifdef CM_TC_TRACE
        long long tr0 = clock64(), tr1, tr2 = 0, tr3;
#endif
        fence_proxy_async();
        fence_before_thread_sync();
        __syncthreads();
#ifdef CM_TC_TRACE
        tr1 = clock64();
#endif
        if (warp == 0) {                 // whole warp, converged: waits for the weights, elected lane issues
            fence_after_thread_sync();
            for (int i = 0; i < nops; ++i) {
                const uint32_t b = consumed + (uint32_t)i;
                ok = mbar_wait(&bars[b & 1u], (b >> 1) & 1u) && ok;
                const TcStage &st = P.st[si + i];
                const MmaOp op = i ? op1 : op0;
                issue_layer(tmem + op.dcol, ACT, WB + (b & 1u) * 32768u, st.N, st.Kp, op.acc);
            }
            if (elect_one()) mma_commit(&bars[2]);
            __syncwarp();
#ifdef CM_TC_TRACE
            tr2 = clock64();
#endif
        }
        ok = mbar_wait(&bars[2], m_phase) && ok;
        m_phase ^= 1;
        fence_after_thread_sync();
#ifdef CM_TC_TRACE
        tr3 = clock64();
        if (blockIdx.x == 0 && tid == 0 && consumed < 64 && io.workspace) {
            long long *tb = reinterpret_cast<long long *>(io.workspace) + 4 * consumed;
            tb[0] = tr0; tb[1] = tr1; tb[2] = tr2; tb[3] = tr3;
        }
#endif
        consumed += (uint32_t)nops;
        si += nops;
        issue_loads();
    };
    const MmaOp none = {0u, 0u};

    // observation panel `pnl` (64 columns, the first one Kp0 wide) of tile `tl` -> registers, 128 * Kp / 512 <= 16 per thread
    float ov[16];
    bool ov_ready = false;
    auto load_obs_panel = [&](float (&dst)[16], int64_t tl, int pnl) {
        const int Kp = P.st[pnl].Kp, kofs = 64 * pnl, total = kTcRows * Kp;
        const int64_t e0 = tl * A.envs_per_tile;
        const int rws = (int)min((int64_t)A.envs_per_tile, io.n_envs - e0) * n;
        const float *src = io.obs + e0 * n * D;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int e = tid + q * kTcThreads;
            const int r = e / Kp, k = e - r * Kp;
            dst[q] = (e < total && r < rws && kofs + k < D) ? __ldg(src + (size_t)r * D + kofs + k) : 0.0f;
        }
    };

    for (int64_t tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x) {
        const int64_t env0 = tile * A.envs_per_tile;
        const int envs = (int)min((int64_t)A.envs_per_tile, io.n_envs - env0);
        const int rows = envs * n;
        const int64_t row0 = env0 * n;
        const bool valid = row < rows;
        const int el = valid ? row / n : 0, il = row - el * n, j0 = el * n;
        const int64_t env = env0 + el, g = row0 + row;
        si = 0;
        // neighbour masks of this row for every layer: fetched now, used after the attention (global latency hidden)
        uint32_t msk0[CM_MAX_LAYERS], msk1[CM_MAX_LAYERS];
#pragma unroll
        for (int l = 0; l < CM_MAX_LAYERS; ++l) {
            uint32_t m0 = 0u, m1 = 0u;
            if (valid && l < L) {
                m0 = m1 = 0xFFFFFFFFu;
                if (io.adj_bits) {
                    const uint32_t *p = io.adj_bits + (env * n + il) * W;
                    m0 &= __ldg(p);
                    if (W > 1) m1 &= __ldg(p + 1);
                }
                if (io.chan_bits) {
                    const uint32_t *p = io.chan_bits + ((env * L + l) * n + il) * W;
                    m0 &= __ldg(p);
                    if (W > 1) m1 &= __ldg(p + 1);
                }
            }
            msk0[l] = m0;
            msk1[l] = m1;
        }

        // ---------------- encoder layer 1: obs panels -> DA[0:128] ----------------
        for (int pnl = 0; pnl < P.l1_panels; ++pnl) {
            const int Kp = P.st[si].Kp;
            const uint32_t lo_off = (uint32_t)kTcRows * Kp * 2;
            // all global loads first (independent, 128 * Kp / 512 <= 16 per thread), then the split + stores; the first
            // panel was already fetched into `ov` while the previous tile was in its head layers
            const int total = kTcRows * Kp;
            if (pnl > 0 || !ov_ready) load_obs_panel(ov, tile, pnl);
            ov_ready = false;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int e = tid + q * kTcThreads;
                if (e < total) {
                    const int r = e / Kp, k = e - r * Kp;
                    __half h, l;
                    split16(ov[q], h, l);
                    const uint32_t off = canon_off16(r, k, Kp);
                    *reinterpret_cast<__half *>(ACT + off) = h;
                    *reinterpret_cast<__half *>(ACT + lo_off + off) = l;
                }
            }
            run_mma(1, MmaOp{DA, (uint32_t)pnl}, none);
        }
        // ---------------- encoder layer 2 (K = 128 as two panels of h) -> DB[0:64] ----------------
        for (int p = 0; p < 2; ++p) {
            float v[16];
            ld_acc<16>(lane_addr, DA, 128, 64 * p + 16 * sub, v);
#pragma unroll
            for (int c = 0; c < 16; ++c) v[c] = tanh_fast(v[c] + bias_s[kBEnc1 + 64 * p + 16 * sub + c]);
            write_act<16>(ACT, 64, row, 16 * sub, v);
            run_mma(1, MmaOp{DB, (uint32_t)p}, none);
        }
        // ---------------- E = tanh(. + b2): k-major fp32 copy + A operand; Q -> DA[0:64], H_0 Wg_0 -> DA[64:128] ----------------
        {
            float v[16];
            ld_acc<16>(lane_addr, DB, 64, 16 * sub, v);
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                v[c] = tanh_fast(v[c] + bias_s[kBEnc2 + 16 * sub + c]);
                ET[(16 * sub + c) * kTPitch + row] = v[c];
            }
            write_act<16>(ACT, 64, row, 16 * sub, v);
            run_mma(2, MmaOp{DA, 0u}, MmaOp{DA + 128, 0u});
        }
        // ---------------- scores, softmax (exact per environment, CUDA cores) ----------------
        {
            float v[16];
            ld_acc<16>(lane_addr, DA, 64, 16 * sub, v);
#pragma unroll
            for (int c = 0; c < 16; ++c) QT[(16 * sub + c) * kTPitch + row] = v[c];
        }
        __syncthreads();
        // scores, row softmax -> M^T; QT is overwritten by M^T once every thread has read it
        switch (n <= 4 ? 1 : (n <= 8 ? 2 : (n <= 16 ? 4 : (n <= 32 ? 8 : 16)))) {
        case 1: scores_softmax<1>(QT, ET, HWT, row, j0, sub, n, valid); break;
        case 2: scores_softmax<2>(QT, ET, HWT, row, j0, sub, n, valid); break;
        case 4: scores_softmax<4>(QT, ET, HWT, row, j0, sub, n, valid); break;
        case 8: scores_softmax<8>(QT, ET, HWT, row, j0, sub, n, valid); break;
        default: scores_softmax<16>(QT, ET, HWT, row, j0, sub, n, valid); break;
        }
        const float *MT = QT;
        if (io.attention) {               // unmasked softmax (agent_infos['attention_weights'])
            float *dst = io.attention + row0 * n;
            for (int e = tid; e < rows * n; e += kTcThreads) {
                const int r = e / n, jl = e - r * n;
                dst[e] = MT[jl * kTPitch + r];
            }
        }
        // H_0 Wg_0 out of tensor memory (the softmax scratch is dead now)
        {
            float v[16];
            ld_acc<16>(lane_addr, DA + 128, 64, 16 * sub, v);
            __syncthreads();
#pragma unroll
            for (int c = 0; c < 16; ++c) HWT[(16 * sub + c) * kTPitch + row] = v[c];
        }
        __syncthreads();
        // ---------------- graph convolutions ----------------
        for (int l = 0; l < L; ++l) {
            // A_l = M * Range * chan_l / (sum + 1e-12); out = A_l (H_l Wg_l)   (comm_base_net.py:101-103, graph_conv_module.py:51-72)
            uint32_t m0 = 0u, m1 = 0u;
#pragma unroll
            for (int q = 0; q < CM_MAX_LAYERS; ++q)
                if (q == l) { m0 = msk0[q]; m1 = msk1[q]; }
            float acc[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) acc[c] = 0.0f;
            float den = 0.0f;
            const int nn = (valid && !(CM_TC_DEBUG & 4)) ? n : 0;
            for (int jj = 0; jj < nn; ++jj) {
                const uint32_t bit = ((jj < 32 ? m0 : m1) >> (jj & 31)) & 1u;
                const float a = bit ? MT[jj * kTPitch + row] : 0.0f;
                den += a;
                const float *hw = HWT + (16 * sub) * kTPitch + j0 + jj;
#pragma unroll
                for (int c = 0; c < 16; ++c) acc[c] = fmaf(a, hw[c * kTPitch], acc[c]);
            }
            const float inv = 1.0f / (den + 1e-12f);
            float v[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                v[c] = tanh_fast(acc[c] * inv + bias_s[kBGcn + l * 64 + 16 * sub + c]);
                if (l + 1 == L && d.residual) v[c] += ET[(16 * sub + c) * kTPitch + row];   // X = E + H_L
            }
            write_act<16>(ACT, 64, row, 16 * sub, v);
            if (l + 1 < L) {
                run_mma(1, MmaOp{DB, 0u}, none);
                float hv[16];
                ld_acc<16>(lane_addr, DB, 64, 16 * sub, hv);
                __syncthreads();          // every thread is done reading HWT of layer l
#pragma unroll
                for (int c = 0; c < 16; ++c) HWT[(16 * sub + c) * kTPitch + row] = hv[c];
                __syncthreads();
            }
        }
        // ---------------- categorical head ----------------
        if (CM_TC_OBS_PREFETCH && tile + gridDim.x < A.n_tiles) {       // next tile's first observation panel: in flight during the head
            load_obs_panel(ov, tile + gridDim.x, 0);
            ov_ready = true;
        }
        run_mma(1, MmaOp{DA, 0u}, none);                                          // 64 -> 128
        for (int p = 0; p < 2; ++p) {                                             // 128 -> 64 as two K panels
            float v[16];
            ld_acc<16>(lane_addr, DA, 128, 64 * p + 16 * sub, v);
#pragma unroll
            for (int c = 0; c < 16; ++c) v[c] = tanh_fast(v[c] + bias_s[kBH1 + 64 * p + 16 * sub + c]);
            write_act<16>(ACT, 64, row, 16 * sub, v);
            run_mma(1, MmaOp{DB, (uint32_t)p}, none);
        }
        {                                                                         // 64 -> 32
            float v[16];
            ld_acc<16>(lane_addr, DB, 64, 16 * sub, v);
#pragma unroll
            for (int c = 0; c < 16; ++c) v[c] = tanh_fast(v[c] + bias_s[kBH2 + 16 * sub + c]);
            write_act<16>(ACT, 64, row, 16 * sub, v);
            run_mma(1, MmaOp{DA, 0u}, none);
        }
        {                                                                         // 32 -> 5 (padded to 16)
            float v[8];
            ld_acc<8>(lane_addr, DA, 32, 8 * sub, v);
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = tanh_fast(v[c] + bias_s[kBH3 + 8 * sub + c]);
            write_act<8>(ACT, 32, row, 8 * sub, v);
            run_mma(1, MmaOp{DB, 0u}, none);
        }
        // ---------------- softmax, availability mask, renormalise, sample ----------------
        if (sub == 0) {
            float lg8[8];
            ld_acc<8>(lane_addr, DB, 16, 0, lg8);
            if (valid) {
                float lg[CM_ACTIONS], pr[CM_ACTIONS];
                float mx = -INFINITY;
#pragma unroll
                for (int a = 0; a < CM_ACTIONS; ++a) { lg[a] = lg8[a] + bias_s[kBH4 + a]; mx = fmaxf(mx, lg[a]); }
                float sum = 0.0f;
#pragma unroll
                for (int a = 0; a < CM_ACTIONS; ++a) { pr[a] = expf(lg[a] - mx); sum += pr[a]; }
                const uint32_t av = io.avail_bits ? io.avail_bits[g] : 0x1Fu;
                float msum = 0.0f;
#pragma unroll
                for (int a = 0; a < CM_ACTIONS; ++a) { pr[a] = ((av >> a) & 1u) ? pr[a] / sum : 0.0f; msum += pr[a]; }
#pragma unroll
                for (int a = 0; a < CM_ACTIONS; ++a) pr[a] = pr[a] / msum;
                if (io.logits) for (int a = 0; a < CM_ACTIONS; ++a) io.logits[g * CM_ACTIONS + a] = lg[a];
                if (io.probs) for (int a = 0; a < CM_ACTIONS; ++a) io.probs[g * CM_ACTIONS + a] = pr[a];
                if (io.actions) {
                    int act;
                    if (d.greedy) {
                        act = 0;
                        for (int a = 1; a < CM_ACTIONS; ++a) if (pr[a] > pr[act]) act = a;
                    } else {
                        float u;
                        if (io.sample_u) u = io.sample_u[g];
                        else {
                            const uint4 blk = philox4x32_10(
                                make_uint4((uint32_t)(d.env_id0 + env), io.tick[env], kStreamAct | (io.episode[env] << 8), (uint32_t)(il >> 2)),
                                make_uint2((uint32_t)d.seed, (uint32_t)(d.seed >> 32)));
                            const uint32_t w = (il & 3) == 0 ? blk.x : ((il & 3) == 1 ? blk.y : ((il & 3) == 2 ? blk.z : blk.w));
                            u = u24(w);
                        }
                        int last = 4;
                        for (int a = 0; a < CM_ACTIONS; ++a) if (pr[a] > 0.0f) last = a;
                        act = -1;
                        float c = 0.0f;
                        for (int a = 0; a < CM_ACTIONS; ++a) { c += pr[a]; if (act < 0 && u < c) act = a; }
                        if (act < 0) act = last;
                    }
                    io.actions[g] = (int8_t)act;
                }
            }
        }
        // the next tile's first MMA must not overwrite accumulators that are still being read
        fence_before_thread_sync();
        __syncthreads();
        fence_after_thread_sync();
    }
    if (!ok && io.error_flag) atomicExch(io.error_flag, (int)CM_ECUDA);
    fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// weight preparation: fp32 k-major blob -> per-stage canonical hi | lo panels
// ------------------------------------------------------------------------------------------------
__global__ void tc_prepare_kernel(const float *__restrict__ w, __half *__restrict__ out, const TcPlan P)
{
    for (int s = 0; s < P.n_stages; ++s) {
        const TcStage st = P.st[s];
        const int total = st.N * st.Kp;
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
            const int r = e / st.Kp, k = e - r * st.Kp;
            float v = 0.0f;
            if (r < st.Nsrc && st.k0 + k < st.Ksrc) v = w[st.src_off + (size_t)(st.k0 + k) * st.Nsrc + r];
            const __half h = __float2half_rn(v);
            const __half l = __float2half_rn((v - __half2float(h)) * 4096.0f);
            // canonical K-major layout over the stacked 2N rows: hi rows [0, N), lo rows [N, 2N)
            auto idx = [&](int rr) { return (rr >> 3) * (st.Kp >> 3) * 64 + (k >> 3) * 64 + (rr & 7) * 8 + (k & 7); };
            out[st.w_off + idx(r)] = h;
            out[st.w_off + idx(st.N + r)] = l;
        }
    }
}

static size_t tc_smem_bytes() { return (size_t)kActBytes + kWBytes + 3 * 64 * kTPitch * 4 + kBiasFloats * 4 + 64; }

int launch_policy_tc(const cm_policy_desc *desc, const cm_policy_io *io, cudaStream_t stream)
{
    if (!io->tc_weights) return CM_EINVAL;
    if (desc->n_agents > 64 || desc->obs_dim > 128) return CM_EUNSUPPORTED;
    TcArgs A;
    A.d = *desc;
    A.io = *io;
    A.plan = make_tc_plan(desc->obs_dim, desc->n_layers);
    A.envs_per_tile = kTcRows / desc->n_agents;
    A.n_tiles = (io->n_envs + A.envs_per_tile - 1) / A.envs_per_tile;
    const size_t smem = tc_smem_bytes();
    static thread_local struct { int dev; int sms; } cache = {-1, 0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return set_cuda_error(cudaGetLastError(), CM_ENODEVICE);
    if (cache.dev != dev) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return set_cuda_error(cudaGetLastError(), CM_ECUDA);
        cudaError_t e = cudaFuncSetAttribute(policy_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_cuda_error(e, CM_ECUDA);
        cache.dev = dev; cache.sms = sms;
    }
    const int grid = (int)(A.n_tiles < cache.sms ? A.n_tiles : cache.sms);   // persistent: one CTA per SM
    policy_tc_kernel<<<grid, kTcThreads, smem, stream>>>(A);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, CM_ECUDA);
    return CM_OK;
}

}  // namespace cm

extern "C" size_t cm_policy_tc_blob_floats(int32_t obs_dim, int32_t n_layers)
{
    return (size_t)(cm::make_tc_plan(obs_dim, n_layers).total_halves + 1) / 2;
}

extern "C" int cm_policy_tc_prepare(const cm_policy_desc *desc, const float *weights, float *tc_weights, cm_stream_t stream)
{
    if (!desc || !weights || !tc_weights) return CM_EINVAL;
    if (desc->n_layers < 1 || desc->n_layers > CM_MAX_LAYERS || desc->obs_dim < 1 || desc->obs_dim > 128) return CM_EUNSUPPORTED;
    if (cm_device_count() < 1) return CM_ENODEVICE;
    const cm::TcPlan P = cm::make_tc_plan(desc->obs_dim, desc->n_layers);
    cm::tc_prepare_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(weights, reinterpret_cast<__half *>(tc_weights), P);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? CM_OK : cm::set_cuda_error(e, CM_ECUDA);
}
