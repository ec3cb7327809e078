// Fused channel-routing step for B200 (sm_100a).
//
// One kernel replaces the ~20 whole-grid numpy passes of
// channels_component.update() (components/channels_base.py:803-962):
//   update_R :1548  update_R_integral :1657  update_channel_discharge :1674
//   update_flow_volume :1951  update_channel_depth :2250  update_trapezoid_Rh
//   :2643  update_free_surface_slope :2583  update_shear_stress/_speed :2620
//   update_friction_factor :2714  manning_formula :3960 / law_of_the_wall :4005
//   update_velocity_on_edges :2851  update_froude_number :2867
//   update_outlet_values :2884  update_peak_values :2908  update_Q_out_integral
//   :2928  update_edge_values :2938  check_flow_depth :3193 / _velocity :3352
// and, when the TFB_CH_SRC_* flags are set, the per-cell source terms of the
// met / snow / evap / infil components evaluated just before it.
//
// The reference's D8 inflow SCATTER (vol[p_k] += dt*Q[w_k], k = 1..8) is done
// as a deterministic parent-side GATHER in the same order (child at SW, W, NW,
// N, NE, E, SE, S), which is bit-identical.
//
// Arithmetic is fp64 with the reference's operation order; the TU is compiled
// with -fmad=false so that a*b+c is rounded twice exactly as numpy does.
// Only x**(2/3), x**(1/3) and log() go through CUDA's libdevice (cbrt / log,
// <= 1-2 ulp from numpy's), everything else (+ - * / sqrt) is IEEE-exact.
#include "tfb_common.cuh"

namespace tfb {

struct ChConsts {
    double g, aval, kappa, law_const;
};

// per-thread reduction slots
enum {
    S_VOL_R = 0, S_VOL_EDGE, S_VOL_P, S_VOL_PR, S_VOL_PS, S_VOL_SM, S_VOL_ET,
    S_VOL_IN, S_NNEG_D, S_NNAN_D, S_NINF_D, S_NNEG_U, S_NNAN_U, S_NINF_U,
    S_NNAN_IN, S_NSLOT   // = 15 sums; slot 15 of TFB_NSUM carries outlet flag
};
constexpr int kPartialStride = 24;
static_assert(kPartialStride == TFB_NSUM, "TFB_NSUM must equal the partial stride");   // 16 sums + dt_min + 4 outlet + pad
constexpr int P_DTMIN = 16, P_QOUT = 17, P_UOUT = 18, P_DOUT = 19, P_FOUT = 20,
              P_HASOUT = 21;

struct Acc {
    double s[S_NSLOT];
    float dt_min;
};

__device__ __forceinline__ unsigned code_at(const tfb_channels_step_args &a, int r, int c) {
    if (c < 0 || c >= a.nx || r < -a.halo || r >= a.ny + a.halo) return 0u;
    return a.code[(int64_t)r * a.nx + c];
}

// source terms of one cell ---------------------------------------------------
struct Src {
    double P_rain, SM, GW, MR, ET, IN;
    bool ET_is_grid;
};

template <bool SRC, bool COMMIT>
__device__ __forceinline__ void eval_sources(const tfb_channels_step_args &a, int r,
                                             int64_t i, double da, Src &s, Acc *acc) {
    const double dt = a.dt;
    s.GW = sg_at(a.GW, i);
    s.MR = sg_at(a.MR, i);
    if (!SRC) {
        s.P_rain = sg_at(a.P_rain, i);
        s.SM = sg_at(a.SM, i);
        s.IN = sg_at(a.IN, i);
        s.ET = sg_at(a.ET, i);
        s.ET_is_grid = (a.ET.ptr != nullptr);
        return;
    }
    // --- met: rain/snow split (met_base.py:1300-1388) ------------------------
    double P_snow = 0.0;
    if (a.flags & TFB_CH_SRC_MET) {
        const double P = sg_at(a.P_rain, i);
        const double Ta = sg_at(a.T_air, i);
        s.P_rain = P * ((Ta > a.T_rain_snow) ? 1.0 : 0.0);
        P_snow = P * ((Ta <= a.T_rain_snow) ? 1.0 : 0.0);
        if (COMMIT) {
            acc->s[S_VOL_P] += (P * da) * dt;
            acc->s[S_VOL_PR] += (s.P_rain * da) * dt;
            acc->s[S_VOL_PS] += (P_snow * da) * dt;
        }
    } else {
        s.P_rain = sg_at(a.P_rain, i);
    }
    // --- snow: degree-day (snow_degree_day.py:293-360, snow_base.py:490-703) --
    if (a.flags & TFB_CH_SRC_SNOW) {
        const double Ta = sg_at(a.T_air, i);
        const double h = __ldg(a.h_swe + i);
        double M = (sg_at(a.c0, i) / 8.64E7) * (Ta - sg_at(a.T0, i));
        M = np_min(M, h / dt);
        M = np_max(M, 0.0);
        s.SM = M;
        if (COMMIT) {
            double hn = h + (P_snow * dt);
            hn = hn - (M * dt);
            hn = np_max(hn, 0.0);
            a.h_swe_out[i] = hn;
            if (a.h_snow) a.h_snow[i] = hn * (a.rho_H2O / a.rho_snow);
            acc->s[S_VOL_SM] += (M * da) * dt;
        }
    } else {
        s.SM = sg_at(a.SM, i);
    }
    // --- evap: Priestley-Taylor (evap_priestley_taylor.py:308-413) ------------
    if (a.flags & TFB_CH_SRC_EVAP) {
        const double Ta = sg_at(a.T_air, i);
        const double Qc = a.K_soil * (sg_at(a.T_surf, i) - a.T_soil_x) / a.soil_x;
        const double Qnet = sg_at(a.Qn_SW, i) + sg_at(a.Qn_LW, i);
        const double Qet = (a.alpha * (0.406 + (0.011 * Ta))) * (Qnet + Qc);
        s.ET = np_max(Qet / 2.5E+9, 0.0);
        s.ET_is_grid = true;
        if (COMMIT) acc->s[S_VOL_ET] += (s.ET * da) * dt;
    } else {
        s.ET = sg_at(a.ET, i);
        s.ET_is_grid = (a.ET.ptr != nullptr);
    }
    // --- infil: Green-Ampt v1 (infil_green_ampt.py:511-578) -------------------
    if (a.flags & TFB_CH_SRC_INFIL) {
        const double Ks = sg_at(a.Ks, i);
        const double P_total = s.P_rain + s.SM;
        const double dK = Ks - sg_at(a.Ki, i);
        const double dq = sg_at(a.qs, i) - sg_at(a.qi, i);
        const double Iv = __ldg(a.I + i);
        const double fc = (((sg_at(a.G, i) * dK) * dq) / Iv) + Ks;
        double IN = np_min(fc, P_total);
        if (P_total < Ks) IN = P_total;                       // infil_base.py:954-995
        if (a.elev.ptr || a.h_table.ptr) {                    // infil_base.py:697-701
            if (sg_at(a.h_table, i) >= sg_at(a.elev, i)) IN = 0.0;
        }
        s.IN = IN;
        if (COMMIT) {
            a.I_out[i] = Iv + (IN * dt);
            acc->s[S_VOL_IN] += (IN * da) * dt;
            if (!isfinite(IN)) acc->s[S_NNAN_IN] += 1.0;
        }
    } else {
        s.IN = sg_at(a.IN, i);
    }
    (void)r;
}

// static geometry of one cell --------------------------------------------------
struct Geo {
    unsigned code;
    double ds, da, width, tan_a, cos_a, angle;
};

__device__ __forceinline__ void load_geo(const tfb_channels_step_args &a, int r, int c,
                                         int64_t i, Geo &g) {
    g.code = a.code[i];
    const float ds32 = ds_of(g.code, __ldg(a.dx + r), __ldg(a.dy + r), __ldg(a.dd + r));
    g.ds = sg_at(a.sinu, i) * (double)ds32;                      // channels_base :1242
    g.da = a.da.ptr ? __ldg(a.da.ptr + r) : a.da.val;
    g.width = sg_at(a.width, i);
    g.tan_a = sg_at(a.tan_angle, i);
    g.cos_a = sg_at(a.cos_angle, i);
    g.angle = sg_at(a.angle, i);
    (void)c;
}

// new volume (before the noflow zeroing) and new depth of cell (r, c) ----------
template <bool SRC, bool COMMIT>
__device__ __forceinline__ void vol_depth(const tfb_channels_step_args &a, int r, int c,
                                          int64_t i, const Geo &g, double vol_old,
                                          double Q_own, double &vol_new, double &d_new,
                                          double &R_out, Acc *acc) {
    const double dt = a.dt;
    Src s;
    eval_sources<SRC, COMMIT>(a, r, i, g.da, s, acc);
    // update_R :1625-1649
    double ET = s.ET;
    if (s.ET_is_grid) {
        if (vol_old < ((ET * g.da) * dt)) {
            ET = 0.0;
            if (COMMIT && (a.flags & TFB_CH_ET_INPLACE) && a.ET_rw) a.ET_rw[i] = 0.0;
        }
    } else {
        ET = 0.0;                       // reference: scalar ET is dropped (:1629)
    }
    const double R = (((s.P_rain + s.SM) + s.GW) + s.MR) - (ET + s.IN);
    R_out = R;
    if (COMMIT && a.is_R_grid) acc->s[S_VOL_R] += (R * g.da) * dt;   // :1666-1670
    // update_flow_volume :2086
    double v = vol_old + ((R * g.da) * dt);
    // parent-side gather, children in reference order (codes 1,2,4,...,128):
    // a child with code g_k sits at the OPPOSITE neighbour position.
    //   code 1 (NE) -> child at SW (r+1,c-1);  2 (E) -> W (r,c-1);
    //   4 (SE) -> NW (r-1,c-1);  8 (S) -> N (r-1,c);  16 (SW) -> NE (r-1,c+1);
    //   32 (W) -> E (r,c+1);  64 (NW) -> SE (r+1,c+1);  128 (N) -> S (r+1,c)
    const int cdr[8] = {1, 0, -1, -1, -1, 0, 1, 1};
    const int cdc[8] = {-1, -1, -1, 0, 1, 1, 1, 0};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int rr = r + cdr[k], cc = c + cdc[k];
        if (code_at(a, rr, cc) == (1u << k)) {
            v += dt * __ldg(a.Q_in + ((int64_t)rr * a.nx + cc));   // :2116-2131
        }
    }
    v -= (Q_own * dt);                                             // :2138
    v = np_max(v, 0.0);                                            // :2145
    vol_new = v;
    // update_channel_depth :2299-2330, :2379
    double d;
    if (g.angle == 0.0) {
        d = v / (g.width * g.ds);
    } else {
        const double denom = 2.0 * g.tan_a;
        double arg = ((2.0 * denom) * v) / g.ds;
        arg += g.width * g.width;
        d = (sqrt(arg) - g.width) / denom;
    }
    d_new = np_max(d, 0.0);
}

template <bool DW, bool LOW, bool DIAG, bool SRC>
__device__ __forceinline__ void route_cell(const tfb_channels_step_args &a,
                                           const ChConsts &k, int r, int c, Acc &acc) {
    const int64_t i = (int64_t)r * a.nx + c;
    Geo g;
    load_geo(a, r, c, i, g);
    const double vol_old = __ldg(a.vol + i);
    const double Q_own = __ldg(a.Q_in + i);
    double v, d, R;
    vol_depth<SRC, true>(a, r, c, i, g, vol_old, Q_own, v, d, R, &acc);
    const bool noflow = (g.code == 0);
    // update_trapezoid_Rh :2668-2703
    const double L2 = d * g.tan_a;
    const double A_wet = d * (g.width + L2);
    double P_wet = g.width + ((2.0 * d) / g.cos_a);
    if (noflow) P_wet = 1.0;
    double Rh = A_wet / P_wet;
    if (noflow) Rh = 0.0;
    // slope: bed (kinematic) or free surface (diffusive) :2583-2618
    const double S_bed = __ldg(a.S_bed + i);
    double S = S_bed;
    if (DW) {
        // depth of the D8 parent AFTER this step's volume update, before the
        // noflow zeroing; noflow cells point at cell (0, 0) of the global grid
        int dr, dc;
        code_offset(g.code, dr, dc);
        int pr = r + dr, pc = c + dc;
        double d_par;
        if (noflow) { pr = -a.row0; pc = 0; }
        const bool have = (pr >= -a.halo && pr < a.ny + a.halo);
        if (have) {
            const int64_t ip = (int64_t)pr * a.nx + pc;
            Geo gp;
            load_geo(a, pr, pc, ip, gp);
            double vp, Rp;
            vol_depth<SRC, false>(a, pr, pc, ip, gp, __ldg(a.vol + ip), __ldg(a.Q_in + ip), vp,
                                  d_par, Rp, nullptr);
        } else {
            d_par = 0.0;   // only reachable for noflow cells of a slab that does
                           // not hold global cell (0,0): diagnostics only
        }
        S = S_bed + ((d - d_par) / g.ds);
    }
    const double S2 = fabs(S);
    const double rough = sg_at(a.rough, i);
    // velocity: kinematic_wave.py:84-127 -> :3960-4003 / :4005-4063
    double u, smooth = 0.0;
    if (!LOW) {
        const double cb = cbrt(Rh);
        u = ((cb * cb) * sqrt(S2)) / rough;
    } else {
        smooth = np_max((k.aval / rough) * d, 1.1);
        u = (k.law_const * sqrt(Rh * S2)) * log(smooth);
    }
    if (noflow) u = 0.0;                                            // :2862
    // outlet sample (pre-zero depth, this step's Q) :2896-2904
    if (r + a.row0 == a.outlet_row && c == a.outlet_col) {
        double f_o = 0.0;
        if (d > 0.0) {
            if (!LOW) f_o = k.g * ((rough * rough) / cbrt(d));
            else { const double q = k.kappa / log(smooth); f_o = q * q; }
        }
        double *p = a.partials + (int64_t)gridDim.x * gridDim.y * kPartialStride;
        p[P_QOUT] = Q_own; p[P_UOUT] = u; p[P_DOUT] = d; p[P_FOUT] = f_o;
        p[P_HASOUT] = 1.0;
    }
    if (DIAG) {
        // :2620-2641, :2714-2788, :2867-2880 (all with the pre-zero depth)
        const double slope_t = DW ? S2 : S_bed;
        const double tau = ((a.rho_H2O * k.g) * d) * slope_t;
        double f = 0.0, fr = 0.0;
        if (d > 0.0) {
            if (!LOW) f = k.g * ((rough * rough) / cbrt(d));
            else { const double q = k.kappa / log(smooth); f = q * q; }
            fr = u / sqrt(k.g * d);
        }
        a.Rh[i] = Rh; a.A_wet[i] = A_wet; a.P_wet[i] = P_wet;
        a.f[i] = f; a.tau[i] = tau; a.u_star[i] = sqrt(tau / a.rho_H2O);
        a.froude[i] = fr;
        if (DW && a.S_free) a.S_free[i] = S;
    }
    if ((a.flags & TFB_CH_WRITE_R) && a.R) a.R[i] = R;
    // update_edge_values :2960-2966
    if (noflow) {
        acc.s[S_VOL_EDGE] += v;
        v = 0.0;
        d = 0.0;
    }
    // checks :3193-3279, :3352-3454
    if (d < 0.0) acc.s[S_NNEG_D] += 1.0;
    if (d != d) acc.s[S_NNAN_D] += 1.0;
    if (isinf(d)) acc.s[S_NINF_D] += 1.0;
    if (u < 0.0) acc.s[S_NNEG_U] += 1.0;
    if (u != u) acc.s[S_NNAN_U] += 1.0;
    if (isinf(u)) acc.s[S_NINF_U] += 1.0;
    if (u > 0.0) acc.dt_min = fminf(acc.dt_min, __fdividef((float)g.ds, (float)u));
    a.vol_out[i] = v;
    a.d[i] = d;
    a.u[i] = u;
    a.Q_out[i] = u * A_wet;                                        // next :1697
}

// fold the block partials (fixed order => deterministic) ------------------------
__device__ void fold_scalars(const tfb_channels_step_args &a, const double *sum,
                             double dt_min, const double *outlet) {
    tfb_channels_scalars &s = *a.scalars;
    const double dt = a.dt;
    if (a.is_R_grid) {
        s.vol_R += sum[S_VOL_R];
    } else {
        // scalar R: (R*da*dt) * n_pixels  (channels_base :1666-1668); the kernel
        // left R in sum[S_VOL_R] slot unused, so recompute from the scalars
        double ET = 0.0;   // scalar ET is dropped by the reference (:1629)
        const double R = (((a.P_rain.val + a.SM.val) + a.GW.val) + a.MR.val) - (ET + a.IN.val);
        s.vol_R += ((R * a.da.val) * dt) * (double)a.n_pixels_global;
    }
    s.vol_edge += sum[S_VOL_EDGE];
    s.vol_P += sum[S_VOL_P];   s.vol_PR += sum[S_VOL_PR]; s.vol_PS += sum[S_VOL_PS];
    s.vol_SM += sum[S_VOL_SM]; s.vol_ET += sum[S_VOL_ET]; s.vol_IN += sum[S_VOL_IN];
    s.n_neg_d = (int64_t)sum[S_NNEG_D]; s.n_nan_d = (int64_t)sum[S_NNAN_D];
    s.n_inf_d = (int64_t)sum[S_NINF_D]; s.n_neg_u = (int64_t)sum[S_NNEG_U];
    s.n_nan_u = (int64_t)sum[S_NNAN_U]; s.n_inf_u = (int64_t)sum[S_NINF_U];
    s.n_nan_IN = (int64_t)sum[S_NNAN_IN];
    s.dt_cfl_min = dt_min;
    if (outlet) {
        s.Q_outlet = outlet[0]; s.u_outlet = outlet[1];
        s.d_outlet = outlet[2]; s.f_outlet = outlet[3];
    }
    // update_peak_values :2908-2926, update_Q_out_integral :2934
    if (s.Q_outlet > s.Q_peak) { s.Q_peak = s.Q_outlet; s.T_peak = a.time_min; }
    if (s.u_outlet > s.u_peak) { s.u_peak = s.u_outlet; s.Tu_peak = a.time_min; }
    if (s.d_outlet > s.d_peak) { s.d_peak = s.d_outlet; s.Td_peak = a.time_min; }
    s.vol_Q += (s.Q_outlet * dt);
    s.step_count += 1;
}

constexpr int kBX = 32, kBY = 8, kRowsPerThread = 2;
constexpr int kTileW = kBX, kTileH = kBY * kRowsPerThread;

template <bool DW, bool LOW, bool DIAG, bool SRC>
__global__ void __launch_bounds__(kBX * kBY)
channels_step_kernel(const tfb_channels_step_args a, const ChConsts k) {
    Acc acc;
#pragma unroll
    for (int j = 0; j < S_NSLOT; ++j) acc.s[j] = 0.0;
    acc.dt_min = 3.0e38f;
    const int c = blockIdx.x * kTileW + threadIdx.x;
    const int rbase = blockIdx.y * kTileH + threadIdx.y;
    if (c < a.nx) {
#pragma unroll
        for (int j = 0; j < kRowsPerThread; ++j) {
            const int r = rbase + j * kBY;
            if (r < a.ny) route_cell<DW, LOW, DIAG, SRC>(a, k, r, c, acc);
        }
    }
    // ---- block reduction: warp shuffles, then one smem hop ------------------
    __shared__ double sm[kBY][S_NSLOT + 1];
    __shared__ bool is_last;
    const int lane = threadIdx.x, warp = threadIdx.y;
#pragma unroll
    for (int j = 0; j < S_NSLOT; ++j) {
        const double v = warp_sum(acc.s[j]);
        if (lane == 0) sm[warp][j] = v;
    }
    {
        const double m = warp_min((double)acc.dt_min);
        if (lane == 0) sm[warp][S_NSLOT] = m;
    }
    __syncthreads();
    const int tid = warp * kBX + lane;
    const int nblk = gridDim.x * gridDim.y;
    const int bid = blockIdx.y * gridDim.x + blockIdx.x;
    if (tid <= S_NSLOT) {
        double v = sm[0][tid];
        for (int w = 1; w < kBY; ++w)
            v = (tid == S_NSLOT) ? fmin(v, sm[w][tid]) : (v + sm[w][tid]);
        a.partials[(int64_t)bid * kPartialStride + tid] = v;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned t = atomicAdd(a.ticket, 1u);
        is_last = (t == (unsigned)nblk - 1u);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // last block: every thread folds a strided subset in a FIXED order
    __shared__ double red[kBX * kBY];
    double tot[S_NSLOT + 1];
    const volatile double *P = a.partials;
    for (int j = 0; j <= S_NSLOT; ++j) {
        double v = (j == S_NSLOT) ? 3.0e38 : 0.0;
        for (int b = tid; b < nblk; b += kBX * kBY) {
            const double x = P[(int64_t)b * kPartialStride + j];
            v = (j == S_NSLOT) ? fmin(v, x) : (v + x);
        }
        red[tid] = v;
        __syncthreads();
        for (int s = (kBX * kBY) / 2; s > 0; s >>= 1) {
            if (tid < s)
                red[tid] = (j == S_NSLOT) ? fmin(red[tid], red[tid + s])
                                          : (red[tid] + red[tid + s]);
            __syncthreads();
        }
        tot[j] = red[0];
        __syncthreads();
    }
    if (tid == 0) {
        double *outp = a.partials + (int64_t)nblk * kPartialStride;
        const bool has_out = (outp[P_HASOUT] != 0.0);
        if (a.finalize) {
            double o4[4] = {outp[P_QOUT], outp[P_UOUT], outp[P_DOUT], outp[P_FOUT]};
            fold_scalars(a, tot, tot[S_NSLOT], has_out ? o4 : nullptr);
        } else {
            // leave per-rank sums at the head of `partials` for the allreduce:
            // [0..15) sums, [15] has-outlet flag, then dt_min + outlet values
            double *h = a.partials + (int64_t)(nblk + 1) * kPartialStride;
            for (int j = 0; j < S_NSLOT; ++j) h[j] = tot[j];
            h[S_NSLOT] = has_out ? 1.0 : 0.0;
            h[P_DTMIN] = tot[S_NSLOT];
            h[P_QOUT] = has_out ? outp[P_QOUT] : 0.0;
            h[P_UOUT] = has_out ? outp[P_UOUT] : 0.0;
            h[P_DOUT] = has_out ? outp[P_DOUT] : 0.0;
            h[P_FOUT] = has_out ? outp[P_FOUT] : 0.0;
        }
        outp[P_HASOUT] = 0.0;
        *a.ticket = 0u;    // re-arm for the next step
    }
}

__global__ void channels_fold_kernel(const tfb_channels_step_args a, const double *sums) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    // sums layout = the `h` block above after an allreduce(SUM) over ranks; the
    // dt_min slot must have been reduced with MIN by the caller (or is ignored)
    const bool has_out = sums[S_NSLOT] != 0.0;
    double o4[4] = {sums[P_QOUT], sums[P_UOUT], sums[P_DOUT], sums[P_FOUT]};
    fold_scalars(a, sums, sums[P_DTMIN], has_out ? o4 : nullptr);
}

template <bool DW, bool LOW, bool DIAG, bool SRC>
static int launch(const tfb_channels_step_args &a, const ChConsts &k, cudaStream_t st) {
    dim3 block(kBX, kBY);
    dim3 grid((a.nx + kTileW - 1) / kTileW, (a.ny + kTileH - 1) / kTileH);
    channels_step_kernel<DW, LOW, DIAG, SRC><<<grid, block, 0, st>>>(a, k);
    TFB_CUDA_CHECK(cudaGetLastError());
    return TFB_OK;
}

}  // namespace tfb

using namespace tfb;

extern "C" int64_t tfb_channels_partials_len(int32_t ny, int32_t nx) {
    const int64_t gx = (nx + kTileW - 1) / kTileW, gy = (ny + kTileH - 1) / kTileH;
    return (gx * gy + 2) * kPartialStride;
}

extern "C" int tfb_channels_step(const tfb_channels_step_args *ap, void *stream) {
    TFB_ARG_CHECK(ap != nullptr);
    const tfb_channels_step_args &a = *ap;
    TFB_ARG_CHECK(a.ny > 0 && a.nx > 0 && a.halo >= 0);
    TFB_ARG_CHECK(a.code && a.dx && a.dy && a.dd && a.S_bed);
    TFB_ARG_CHECK(a.vol && a.vol_out && a.Q_in && a.Q_out && a.d && a.u);
    TFB_ARG_CHECK(a.Q_in != a.Q_out);
    TFB_ARG_CHECK(a.scalars && a.partials && a.ticket);
    TFB_ARG_CHECK(a.dt > 0.0);
    const bool dw = a.flags & TFB_CH_DIFFUSIVE, low = a.flags & TFB_CH_LAW_OF_WALL;
    const bool diag = a.flags & TFB_CH_DIAG;
    const bool src = a.flags & (TFB_CH_SRC_MET | TFB_CH_SRC_SNOW | TFB_CH_SRC_EVAP |
                                TFB_CH_SRC_INFIL);
    if (diag)
        TFB_ARG_CHECK(a.Rh && a.A_wet && a.P_wet && a.f && a.tau && a.u_star && a.froude);
    if (a.flags & TFB_CH_SRC_SNOW) TFB_ARG_CHECK(a.h_swe && a.h_swe_out);
    if (a.flags & TFB_CH_SRC_INFIL) TFB_ARG_CHECK(a.I && a.I_out);
    if (a.flags & TFB_CH_DIFFUSIVE) {
        // neighbours re-read pre-step state: it must not be updated in place
        TFB_ARG_CHECK(a.vol != a.vol_out);
        if (a.flags & TFB_CH_SRC_SNOW) TFB_ARG_CHECK(a.h_swe != a.h_swe_out);
        if (a.flags & TFB_CH_SRC_INFIL) TFB_ARG_CHECK(a.I != a.I_out);
    }
    if (src) TFB_ARG_CHECK(a.is_R_grid == 1);
    ChConsts k;
    k.g = 9.81; k.aval = 0.476; k.kappa = 0.408;
    k.law_const = sqrt(k.g) / k.kappa;                 // channels_base :531
    cudaStream_t st = (cudaStream_t)stream;
#define TFB_DISPATCH(DW_, LOW_, DIAG_, SRC_)                                       \
    if (dw == DW_ && low == LOW_ && diag == DIAG_ && src == SRC_)                  \
        return launch<DW_, LOW_, DIAG_, SRC_>(a, k, st);
    TFB_DISPATCH(false, false, false, false) TFB_DISPATCH(false, false, false, true)
    TFB_DISPATCH(false, false, true, false)  TFB_DISPATCH(false, false, true, true)
    TFB_DISPATCH(false, true, false, false)  TFB_DISPATCH(false, true, false, true)
    TFB_DISPATCH(false, true, true, false)   TFB_DISPATCH(false, true, true, true)
    TFB_DISPATCH(true, false, false, false)  TFB_DISPATCH(true, false, false, true)
    TFB_DISPATCH(true, false, true, false)   TFB_DISPATCH(true, false, true, true)
    TFB_DISPATCH(true, true, false, false)   TFB_DISPATCH(true, true, false, true)
    TFB_DISPATCH(true, true, true, false)    TFB_DISPATCH(true, true, true, true)
#undef TFB_DISPATCH
    return TFB_ERR_ARG;
}

extern "C" int tfb_channels_fold(const tfb_channels_step_args *ap, const double *sums,
                                 void *stream) {
    TFB_ARG_CHECK(ap != nullptr && sums != nullptr && ap->scalars != nullptr);
    channels_fold_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(*ap, sums);
    TFB_CUDA_CHECK(cudaGetLastError());
    return TFB_OK;
}
