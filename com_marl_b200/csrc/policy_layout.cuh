// policy_layout.cuh — weight-blob layouts shared by the policy kernels (fp32 FFMA and tcgen05 variants).
#pragma once
#include <stdint.h>

#include "commarl_b200.h"

namespace cm {

static constexpr int kH1 = 128, kE = 64, kC1 = 128, kC2 = 64, kC3 = 32;

struct Blob {   // float offsets into the weight blob (commarl_b200.h)
    int enc_w1, enc_b1, enc_w2, enc_b2, att_w, gcn_w, gcn_b, head_w1, head_b1, head_w2, head_b2, head_w3, head_b3,
        head_w4, head_b4, total;
};

__host__ __device__ inline Blob blob_layout(int D, int L)
{
    Blob o;
    int p = 0;
    o.enc_w1 = p; p += D * kH1;
    o.enc_b1 = p; p += kH1;
    o.enc_w2 = p; p += kH1 * kE;
    o.enc_b2 = p; p += kE;
    o.att_w = p; p += kE * kE;
    o.gcn_w = p; p += L * kE * kE;
    o.gcn_b = p; p += L * kE;
    o.head_w1 = p; p += kE * kC1;
    o.head_b1 = p; p += kC1;
    o.head_w2 = p; p += kC1 * kC2;
    o.head_b2 = p; p += kC2;
    o.head_w3 = p; p += kC2 * kC3;
    o.head_b3 = p; p += kC3;
    o.head_w4 = p; p += kC3 * CM_ACTIONS;
    o.head_b4 = p; p += CM_ACTIONS;
    o.total = p;
    return o;
}

struct PolicyArgs {
    cm_policy_desc d;
    cm_policy_io io;
    int envs_per_tile;
    int64_t n_tiles;
};


// ---- tcgen05 variant: one "stage" = the B operand of one tcgen05 product, N rows x Kp columns, stored in the
// canonical K-major core-matrix layout as a hi block followed by a lo block (error-compensated TF32).  Every
// stage is at most 32 KB so that two of them fit the kernel's weight ring; wide layers (N = 128 with Kp = 64)
// are split into two N = 64 stages.  `seq` is the order in which one tile consumes the stages. ----
struct TcStage { int w_off, src_off, N, Kp, k0, n0, Ksrc, Nsrc; };
struct TcPlan {
    int n_stages, total_floats, seq_len;
    int l1_panels, l1_split, h1_split;      // obs K panels (1 or 2); L1 / head-1 issued as two N = 64 halves?
    int bias_floats;
    TcStage st[24];
};

__host__ __device__ inline TcPlan make_tc_plan(int D, int L)
{
    const Blob o = blob_layout(D, L);
    TcPlan P;
    int s = 0, off = 0;
    auto add = [&](int src_off, int N, int Kp, int k0, int n0, int Ksrc, int Nsrc) {
        P.st[s].w_off = off; P.st[s].src_off = src_off; P.st[s].N = N; P.st[s].Kp = Kp; P.st[s].k0 = k0; P.st[s].n0 = n0;
        P.st[s].Ksrc = Ksrc; P.st[s].Nsrc = Nsrc;
        off += 2 * N * Kp;
        return s++;
    };
    // a wide (N = 128) product is one stage when it fits 32 KB, otherwise two N = 64 halves
    auto add_wide = [&](int src_off, int Kp, int k0, int Ksrc) {
        if (2 * 128 * Kp * 4 <= 32768) { add(src_off, 128, Kp, k0, 0, Ksrc, 128); return 0; }
        add(src_off, 64, Kp, k0, 0, Ksrc, 128);
        add(src_off, 64, Kp, k0, 64, Ksrc, 128);
        return 1;
    };
    const int Dp = (D + 7) / 8 * 8;
    P.l1_panels = Dp > 64 ? 2 : 1;
    P.l1_split = add_wide(o.enc_w1, Dp <= 64 ? Dp : 64, 0, D);
    if (Dp > 64) add(o.enc_w1, kH1, Dp - 64, 64, 0, D, kH1);            // second K panel: at most 128 x 64 ... (Dp-64 <= 64)
    add(o.enc_w2, kE, 64, 0, 0, kH1, kE);
    add(o.enc_w2, kE, 64, 64, 0, kH1, kE);
    add(o.att_w, kE, 64, 0, 0, kE, kE);
    for (int l = 0; l < L; ++l) add(o.gcn_w + l * kE * kE, kE, 64, 0, 0, kE, kE);
    P.h1_split = add_wide(o.head_w1, 64, 0, kE);
    add(o.head_w2, kC2, 64, 0, 0, kC1, kC2);
    add(o.head_w2, kC2, 64, 64, 0, kC1, kC2);
    add(o.head_w3, kC3, 64, 0, 0, kC2, kC3);
    add(o.head_w4, 16, 32, 0, 0, kC3, CM_ACTIONS);                       // 5 logits padded to N = 16
    P.n_stages = s;
    P.seq_len = s;                                                        // stages are stored in consumption order
    P.total_floats = off;
    P.bias_floats = o.total - o.enc_b1;                                  // upper bound, biases are read from the fp32 blob
    return P;
}

}  // namespace cm
