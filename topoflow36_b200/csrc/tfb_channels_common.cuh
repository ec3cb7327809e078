// Definitions shared by the generic and the packed routing kernels: reduction
// slots, per-thread accumulators and the fold of reduced sums into the scalar
// block (update_R_integral :1657, update_outlet_values :2884, update_peak_values
// :2908, update_Q_out_integral :2928, update_edge_values :2938 of
// components/channels_base.py).
#pragma once
#include "tfb_common.cuh"

namespace tfb {

struct ChConsts {
    double g, aval, kappa, law_const;
    double R_uniform, Rdadt_uniform;   // UR kernels: all six source inputs scalar
};

// reduction slots ------------------------------------------------------------
enum {
    S_VOL_R = 0, S_VOL_EDGE, S_VOL_P, S_VOL_PR, S_VOL_PS, S_VOL_SM, S_VOL_ET,
    S_VOL_IN, S_NNEG_D, S_NNAN_D, S_NINF_D, S_NNEG_U, S_NNAN_U, S_NINF_U,
    S_NNAN_IN, S_NSLOT   // = 15
};
constexpr int kNSum = 8;        // fp64 sums  (slots 0..7)
constexpr int kNCnt = 7;        // counters   (slots 8..14)
constexpr int kPartialStride = 24;
static_assert(kPartialStride == TFB_NSUM, "TFB_NSUM must equal the partial stride");
constexpr int P_HASOUT_SUM = 15, P_DTMIN = 16, P_QOUT = 17, P_UOUT = 18, P_DOUT = 19,
              P_FOUT = 20, P_HASOUT = 21, P_FAIL = 22;

// (only slots 0..1 are live unless source terms are fused; the compiler drops
// the dead registers because every use is guarded by a template constant)
struct Acc {
    double s[kNSum];
    unsigned n[kNCnt];
    float dt_min;
};

// fold the (locally or globally) reduced sums into the scalar block -------------
// (a) the volume integrals
__device__ inline void fold_sums(const tfb_channels_step_args &a, const double *h) {
    tfb_channels_scalars &s = *a.scalars;
    const double dt = a.dt;
    if (a.is_R_grid) {
        s.vol_R += h[S_VOL_R];
    } else {
        // scalar R: (R*da*dt) * n_pixels (channels_base :1666-1668); a scalar ET
        // is dropped by the reference (:1629)
        const double R = (((a.P_rain.val + a.SM.val) + a.GW.val) + a.MR.val) - (0.0 + a.IN.val);
        s.vol_R += ((R * a.da.val) * dt) * (double)a.n_pixels_global;
    }
    s.vol_edge += h[S_VOL_EDGE];
    s.vol_P += h[S_VOL_P];   s.vol_PR += h[S_VOL_PR]; s.vol_PS += h[S_VOL_PS];
    s.vol_SM += h[S_VOL_SM]; s.vol_ET += h[S_VOL_ET]; s.vol_IN += h[S_VOL_IN];
    s.n_nan_IN = (int64_t)h[S_NNAN_IN];
}
// (b) stability counters, outlet sample, peaks, outlet integral, step counter
__device__ inline void fold_status(const tfb_channels_step_args &a, const double *h) {
    tfb_channels_scalars &s = *a.scalars;
    s.n_neg_d = (int64_t)h[S_NNEG_D]; s.n_nan_d = (int64_t)h[S_NNAN_D];
    s.n_inf_d = (int64_t)h[S_NINF_D]; s.n_neg_u = (int64_t)h[S_NNEG_U];
    s.n_nan_u = (int64_t)h[S_NNAN_U]; s.n_inf_u = (int64_t)h[S_NINF_U];
    s.dt_cfl_min = h[P_DTMIN];
    if (h[P_HASOUT_SUM] != 0.0) {
        s.Q_outlet = h[P_QOUT]; s.u_outlet = h[P_UOUT];
        s.d_outlet = h[P_DOUT]; s.f_outlet = h[P_FOUT];
    }
    // update_peak_values :2908-2926, update_Q_out_integral :2934
    if (s.Q_outlet > s.Q_peak) { s.Q_peak = s.Q_outlet; s.T_peak = a.time_min; }
    if (s.u_outlet > s.u_peak) { s.u_peak = s.u_outlet; s.Tu_peak = a.time_min; }
    if (s.d_outlet > s.d_peak) { s.d_peak = s.d_outlet; s.Td_peak = a.time_min; }
    s.vol_Q += (s.Q_outlet * a.dt);
    s.step_count += 1;
}
// (c) CTAs that gave up a bounded wait (TMA pipeline, peer halo): the step is invalid.
// Cumulative, so the host sees it whenever it next reads the block.
__device__ inline void fold_fail(const tfb_channels_step_args &a, const double *h) {
    a.scalars->n_step_failed += (int64_t)h[P_FAIL];
}
// copy the scalar block to the pinned-host mirror, if any (called by the folding thread)
__device__ inline void mirror_scalars(const tfb_channels_step_args &a) {
    if (!a.scalars_mirror) return;
    const double *src = reinterpret_cast<const double *>(a.scalars);
    volatile double *dst = reinterpret_cast<volatile double *>(a.scalars_mirror);
#pragma unroll 4
    for (int j = 0; j < (int)(sizeof(tfb_channels_scalars) / sizeof(double)); ++j) dst[j] = src[j];
    __threadfence_system();
}
__device__ inline void fold_scalars(const tfb_channels_step_args &a, const double *h) {
    fold_sums(a, h);
    fold_status(a, h);
    fold_fail(a, h);
    mirror_scalars(a);
}

}  // namespace tfb
