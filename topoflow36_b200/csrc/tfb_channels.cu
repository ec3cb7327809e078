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
// D8 inflow: the reference SCATTERS (vol[p_k] += dt*Q[w_k], k = 1..8).  Here each
// cell GATHERS from its children in the same order (child at SW, W, NW, N, NE,
// E, SE, S = codes 1, 2, 4, ..., 128), which is bit-identical.
// Which neighbours are children is precomputed once into a per-cell bit mask
// (tfb_channels_pack_geo), so the hot loop does 1 two-byte load instead of 8
// neighbour-code loads and compares.
//
// Mapping: a thread owns VEC (=2 when nx is even) adjacent cells of a row and
// moves them with 128-bit loads/stores; a warp covers 64 columns of one row, a
// CTA 8 rows; the grid is persistent (a multiple of the SM count) and strides
// over 64 x 64 tiles, so neighbour rows are L1/L2 hits and only one row of
// block partials per CTA reaches the deterministic last-block fold.
//
// Arithmetic is fp64 in the reference's operation order; the TU is compiled
// with -fmad=false so a*b+c rounds twice exactly as numpy does.  Only
// x**(2/3), x**(1/3) and log() go through libdevice (cbrt / log, <= 2 ulp from
// numpy's pow/log); + - * / sqrt are IEEE-exact.
#include "tfb_common.cuh"
#include "tfb_channels_common.cuh"

// resident CTAs (of 256 threads) per SM the register allocator must allow:
// the lean kinematic kernel wants occupancy more than registers (measured)
#ifndef TFB_KW_MINBLOCKS
#define TFB_KW_MINBLOCKS 3
#endif
#ifndef TFB_OTHER_MINBLOCKS
#define TFB_OTHER_MINBLOCKS 2
#endif

namespace tfb {

// child position for mask bit k (= child's code 1<<k): opposite of direction k
__device__ __constant__ const int8_t kChildDR[8] = {1, 0, -1, -1, -1, 0, 1, 1};
__device__ __constant__ const int8_t kChildDC[8] = {-1, -1, -1, 0, 1, 1, 1, 0};

// ---------------------------------------------------------------------------
// geo word: low byte = D8 code, high byte = child mask (bit k set <=> the
// neighbour at the child position of code 1<<k has exactly that code)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_geo_kernel(const uint8_t *__restrict__ code, int ny, int nx, int halo,
                uint16_t *__restrict__ geo) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = (int)(blockIdx.y * blockDim.y + threadIdx.y) - halo;
    if (c >= nx || r >= ny + halo) return;
    const int64_t i = (int64_t)r * nx + c;
    unsigned m = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int rr = r + kChildDR[k], cc = c + kChildDC[k];
        if (cc < 0 || cc >= nx || rr < -halo || rr >= ny + halo) continue;
        if (code[(int64_t)rr * nx + cc] == (uint8_t)(1u << k)) m |= (1u << k);
    }
    geo[i] = (uint16_t)(code[i] | (m << 8));
}

// ---------------------------------------------------------------------------
// per-cell inputs
// ---------------------------------------------------------------------------
struct Cell {
    unsigned code, cmask;
    double ds, da, width, tan_a, cos_a;
    double vol, Q;
    double I, h_swe;          // fused-source state in (read by the caller)
    double vb;                // bankfull channel volume (FLOOD_OPTION)
};
struct SrcState {             // fused-source state out (written by the caller)
    double I, h_swe;
};

struct Src {
    double P_rain, SM, GW, MR, ET, IN;
    bool ET_is_grid;
};

// source terms of one cell; COMMIT=false is the diffusive wave re-deriving its
// parent's inflow (no state update, no integrals)
template <bool SRC, bool COMMIT>
__device__ __forceinline__ void eval_sources(const tfb_channels_step_args &a, int64_t i,
                                             const Cell &g, Src &s, SrcState &ns, Acc *acc) {
    const double da = g.da;
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
        const double h = g.h_swe;
        double M = (sg_at(a.c0, i) / 8.64E7) * (Ta - sg_at(a.T0, i));
        M = np_min(M, h / dt);
        M = np_max(M, 0.0);
        s.SM = M;
        if (COMMIT) {
            double hn = h + (P_snow * dt);
            hn = hn - (M * dt);
            hn = np_max(hn, 0.0);
            ns.h_swe = hn;
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
        const double Iv = g.I;
        const double fc = (((sg_at(a.G, i) * dK) * dq) / Iv) + Ks;
        double IN = np_min(fc, P_total);
        if (P_total < Ks) IN = P_total;                       // infil_base.py:954-995
        if (a.elev.ptr || a.h_table.ptr) {                    // infil_base.py:697-701
            if (sg_at(a.h_table, i) >= sg_at(a.elev, i)) IN = 0.0;
        }
        s.IN = IN;
        if (COMMIT) {
            ns.I = Iv + (IN * dt);
            acc->s[S_VOL_IN] += (IN * da) * dt;
            if (!isfinite(IN)) acc->n[S_NNAN_IN - kNSum] += 1u;
        }
    } else {
        s.IN = sg_at(a.IN, i);
    }
}

__device__ __forceinline__ void load_cell_scalar(const tfb_channels_step_args &a, int r,
                                                 int64_t i, Cell &g) {
    const unsigned w = a.geo[i];
    g.code = w & 0xffu;
    g.cmask = w >> 8;
    const float ds32 = ds_of(g.code, __ldg(a.dx + r), __ldg(a.dy + r), __ldg(a.dd + r));
    g.ds = sg_at(a.sinu, i) * (double)ds32;                      // channels_base :1242
    g.da = a.da.ptr ? __ldg(a.da.ptr + r) : a.da.val;
    g.width = sg_at(a.width, i);
    g.tan_a = sg_at(a.tan_angle, i);
    g.cos_a = sg_at(a.cos_angle, i);
    g.vol = __ldg(a.vol + i);
    g.Q = __ldg(a.Q_in + i);
    g.I = (a.flags & TFB_CH_SRC_INFIL) ? __ldg(a.I + i) : 0.0;
    g.h_swe = (a.flags & TFB_CH_SRC_SNOW) ? __ldg(a.h_swe + i) : 0.0;
    g.vb = (a.flags & TFB_CH_FLOOD) ? sg_at(a.vol_bankfull, i) : 0.0;
}

// volume that reaches the channel this step: all of the runoff volume, or -- with
// ATTENUATE -- the outflow of the cell's linear reservoir (channels_base.py:2077-2084)
template <bool COMMIT>
__device__ __forceinline__ double attenuate(const tfb_channels_step_args &a, int64_t i,
                                            double Rv) {
    if (!(a.flags & TFB_CH_ATTENUATE)) return Rv;
    double vs = __ldg(a.vol_stored + i) + Rv;
    const double Q_sides = vs / a.t_drain;
    const double vol_sides = np_min(Q_sides * a.dt, vs);
    if (COMMIT) {
        vs = vs - vol_sides;
        a.vol_stored_out[i] = np_max(vs, 0.0);
    }
    return vol_sides;
}

// new volume (before the noflow zeroing) and new depth of a cell ---------------
template <bool SRC, bool UR, bool COMMIT>
__device__ __forceinline__ void vol_depth(const tfb_channels_step_args &a, const ChConsts &kc,
                                          int64_t i, const Cell &g,
                                          const double *__restrict__ Qc, double &vol_new,
                                          double &d_new, double &df_new, double &R_out,
                                          SrcState &ns, Acc *acc) {
    const double dt = a.dt;
    double v;
    if (UR) {
        // every source input is a scalar (and so is da): R and (R*da)*dt are the
        // same for all cells -- evaluated once on the host in this very order
        R_out = kc.R_uniform;
        v = g.vol + attenuate<COMMIT>(a, i, kc.Rdadt_uniform);       // :2077-2086
    } else {
        Src s;
        eval_sources<SRC, COMMIT>(a, i, g, s, ns, acc);
        // update_R :1625-1649
        double ET = s.ET;
        if (s.ET_is_grid) {
            if (g.vol < ((ET * g.da) * dt)) {
                ET = 0.0;
                if (COMMIT && (a.flags & TFB_CH_ET_INPLACE) && a.ET_rw) a.ET_rw[i] = 0.0;
            }
        } else {
            ET = 0.0;                   // reference: a scalar ET is dropped (:1629)
        }
        const double R = (((s.P_rain + s.SM) + s.GW) + s.MR) - (ET + s.IN);
        R_out = R;
        if (COMMIT && a.is_R_grid) acc->s[S_VOL_R] += (R * g.da) * dt;   // :1666-1670
        v = g.vol + attenuate<COMMIT>(a, i, (R * g.da) * dt);        // :2077-2086
    }
    // children in the reference's order (codes 1,2,4,...,128) :2116-2131;
    // Qc[k] is the child's discharge, meaningful only where the mask bit is set
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (g.cmask & (1u << k)) v += dt * Qc[k];
    }
    v -= (g.Q * dt);                                                 // :2138
    v = np_max(v, 0.0);                                              // :2145
    vol_new = v;
    df_new = 0.0;
    if (a.flags & TFB_CH_FLOOD) {
        // update_channel_volume :2149, update_flood_volume :2158-2185,
        // update_flood_depth :2393-2448: the depth below comes from min(vol, bankfull)
        const double vf = np_max(v - g.vb, 0.0);
        v = np_min(v, g.vb);
        df_new = np_max(vf / ((a.width_ratio * g.width) * g.ds), 0.0);
    }
    // update_channel_depth :2299-2330, :2379  (angle == 0  <=>  tan(angle) == 0)
    double d;
    if (g.tan_a == 0.0) {
        d = v / (g.width * g.ds);
    } else {
        const double denom = 2.0 * g.tan_a;
        double arg = ((2.0 * denom) * v) / g.ds;
        arg += g.width * g.width;
        d = (sqrt(arg) - g.width) / denom;
    }
    d_new = np_max(d, 0.0);
}

// predicated loads of the children's discharge
__device__ __forceinline__ void load_children(const double *__restrict__ Qp, int nx /* row pitch */,
                                              unsigned cmask, double *Qc) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        Qc[k] = 0.0;
        if (cmask & (1u << k)) Qc[k] = __ldg(Qp + (kChildDR[k] * nx + kChildDC[k]));
    }
}

struct CellOut {
    double vol, d, u, Qn;
    double Rh, A_wet, P_wet, f, tau, u_star, froude, S, R;
    double df, vf;            // FLOOD_OPTION: flood depth and volume above bankfull
};

// everything after the depth: Rh, slope, velocity, diagnostics, bookkeeping
template <bool DW, bool LOW, bool DIAG, bool SRC, bool UR>
__device__ __forceinline__ void finish_cell(const tfb_channels_step_args &a,
                                            const ChConsts &k, int r, int c, int64_t i,
                                            const Cell &g, double v, double d, double df,
                                            double slope_in, double rough, Acc &acc,
                                            CellOut &o) {
    const bool noflow = (g.code == 0);
    // update_trapezoid_Rh :2668-2703
    const double L2 = d * g.tan_a;
    const double A_wet = d * (g.width + L2);
    double P_wet = g.width + ((2.0 * d) / g.cos_a);
    if (noflow) P_wet = 1.0;
    double Rh = A_wet / P_wet;
    if (noflow) Rh = 0.0;
    // slope_in: sqrt(|S_bed|) precomputed on the host for the lean kinematic
    // Manning case, S_bed otherwise (diffusive, law of the wall, DIAG)
    double S = slope_in, sqrtS;
    if (DW) {
        // depth of the D8 parent AFTER this step's volume update and before the
        // noflow zeroing (:2606-2607); noflow cells point at global cell (0, 0)
        int dr, dc;
        code_offset(g.code, dr, dc);
        int pr = r + dr, pc = c + dc;
        if (noflow) { pr = -a.row0; pc = 0; }
        double d_par = 0.0;
        if (pr >= -a.halo && pr < a.ny + a.halo) {
            const int64_t ip = (int64_t)pr * a.pitch + pc;
            Cell gp;
            load_cell_scalar(a, pr, ip, gp);
            double Qc[8], vp, Rp, df_par;
            SrcState nsp;
            load_children(a.Q_in + ip, a.pitch, gp.cmask, Qc);
            vol_depth<SRC, UR, false>(a, k, ip, gp, Qc, vp, d_par, df_par, Rp, nsp, nullptr);
            if (a.flags & TFB_CH_FLOOD) d_par = d_par + df_par;      // total depth :2593-2596
        }
        const double d_tot = (a.flags & TFB_CH_FLOOD) ? (d + df) : d;
        S = slope_in + ((d_tot - d_par) / g.ds);
        sqrtS = sqrt(fabs(S));
    } else {
        sqrtS = (DIAG || LOW) ? (LOW ? 0.0 : sqrt(fabs(slope_in))) : slope_in;
    }
    // velocity: kinematic_wave.py:84-127 -> :3960-4003 / :4005-4063
    double u, smooth = 0.0;
    if (!LOW) {
        const double cb = cbrt(Rh);
        u = ((cb * cb) * sqrtS) / rough;
    } else {
        smooth = np_max((k.aval / rough) * d, 1.1);
        const double S2 = DW ? fabs(S) : fabs(slope_in);
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
        double *p = a.partials + kPartialStride;          // row 1: outlet staging
        p[P_QOUT] = g.Q; p[P_UOUT] = u; p[P_DOUT] = d; p[P_FOUT] = f_o;
        p[P_HASOUT] = 1.0;
    }
    if (DIAG) {
        // :2620-2641, :2714-2788, :2867-2880 (all with the pre-zero depth)
        const double slope_t = DW ? fabs(S) : slope_in;
        const double tau = ((a.rho_H2O * k.g) * d) * slope_t;
        double f = 0.0, fr = 0.0;
        if (d > 0.0) {
            if (!LOW) f = k.g * ((rough * rough) / cbrt(d));
            else { const double q = k.kappa / log(smooth); f = q * q; }
            fr = u / sqrt(k.g * d);
        }
        o.Rh = Rh; o.A_wet = A_wet; o.P_wet = P_wet; o.f = f; o.tau = tau;
        o.u_star = sqrt(tau / a.rho_H2O); o.froude = fr; o.S = S;
    }
    // update_edge_values :2960-2966
    if (noflow) {
        acc.s[S_VOL_EDGE] += v;
        v = 0.0;
        d = 0.0;
    }
    // checks :3193-3279, :3352-3454
    if (d < 0.0) acc.n[S_NNEG_D - kNSum] += 1u;
    if (d != d) acc.n[S_NNAN_D - kNSum] += 1u;
    if (isinf(d)) acc.n[S_NINF_D - kNSum] += 1u;
    if (u < 0.0) acc.n[S_NNEG_U - kNSum] += 1u;
    if (u != u) acc.n[S_NNAN_U - kNSum] += 1u;
    if (isinf(u)) acc.n[S_NINF_U - kNSum] += 1u;
    if (u > 0.0) acc.dt_min = fminf(acc.dt_min, __fdividef((float)g.ds, (float)u));
    o.vol = v; o.d = d; o.u = u;
    o.Qn = u * A_wet;                                               // next :1697
    o.df = 0.0; o.vf = 0.0;
    if (a.flags & TFB_CH_FLOOD) {
        // update_flood_discharge :1701-1750 and update_discharge :1791-1862, evaluated
        // here for the NEXT step: d_flood after the noflow zeroing (:2968-2970), S_bed
        // (kinematic) or this step's S_free (diffusive)
        const double dfz = noflow ? 0.0 : df;
        const double cbf = cbrt(dfz);
        const double sq = DW ? sqrt(fabs(S)) : ((DIAG || LOW) ? sqrt(fabs(slope_in)) : slope_in);
        const double uf = ((cbf * cbf) * sq) / a.flood_manning_n;
        const double Af = (a.width_ratio * g.width) * dfz;
        o.Qn = o.Qn + (uf * Af);
        o.df = dfz;
        o.vf = noflow ? 0.0 : np_max(v - g.vb, 0.0);
    }
    (void)i;
}

// ---------------------------------------------------------------------------
// vector helpers: VEC adjacent cells per thread
// ---------------------------------------------------------------------------
template <int VEC> struct V { double v[VEC]; };

template <int VEC>
__device__ __forceinline__ V<VEC> ldv(const double *__restrict__ p, int64_t i) {
    V<VEC> o;
    if (VEC == 2) {
        const double2 t = __ldg(reinterpret_cast<const double2 *>(p + i));
        o.v[0] = t.x; o.v[VEC - 1] = t.y;
    } else {
        o.v[0] = __ldg(p + i);
    }
    return o;
}
template <int VEC>
__device__ __forceinline__ V<VEC> ldsg(const tfb_sg &s, int64_t i) {
    if (s.ptr) return ldv<VEC>(s.ptr, i);
    V<VEC> o;
#pragma unroll
    for (int j = 0; j < VEC; ++j) o.v[j] = s.val;
    return o;
}
template <int VEC>
__device__ __forceinline__ void stv(double *__restrict__ p, int64_t i, const V<VEC> &x) {
    if (VEC == 2) {
        *reinterpret_cast<double2 *>(p + i) = make_double2(x.v[0], x.v[VEC - 1]);
    } else {
        p[i] = x.v[0];
    }
}

constexpr int kBX = 32, kBY = 8;            // warp = one row segment, 8 rows per CTA
constexpr int kTileRows = 64;               // rows a CTA walks per tile

// GP: width / tan / cos / roughness are all grids and sinu, da scalars (the usual
// case) -> straight 128-bit loads, no scalar-or-grid selects.  UR: see vol_depth.
template <int VEC, bool DW, bool LOW, bool DIAG, bool SRC, bool GP, bool UR>
__global__ void __launch_bounds__(kBX * kBY, (DW || DIAG || SRC) ? TFB_OTHER_MINBLOCKS : TFB_KW_MINBLOCKS)
channels_step_kernel(const tfb_channels_step_args a, const ChConsts k) {
    Acc acc;
#pragma unroll
    for (int j = 0; j < kNSum; ++j) acc.s[j] = 0.0;
#pragma unroll
    for (int j = 0; j < kNCnt; ++j) acc.n[j] = 0u;
    acc.dt_min = 3.0e38f;

    constexpr int kTileW = kBX * VEC;
    const int tiles_x = (a.nx + kTileW - 1) / kTileW;
    const int tiles_y = (a.ny + kTileRows - 1) / kTileRows;
    const int n_tiles = tiles_x * tiles_y;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int ty = t / tiles_x, tx = t - ty * tiles_x;
        const int c0 = tx * kTileW + threadIdx.x * VEC;
        if (c0 >= a.nx) continue;
        const int r_end = min(a.ny, (ty + 1) * kTileRows);
        for (int r = ty * kTileRows + threadIdx.y; r < r_end; r += kBY) {
            const int64_t i = (int64_t)r * a.pitch + c0;
            // ---- own inputs, 128-bit where possible -----------------------
            unsigned gw[VEC];
            if (VEC == 2) {
                const unsigned w2 = __ldg(reinterpret_cast<const unsigned *>(a.geo + i));
                gw[0] = w2 & 0xffffu; gw[VEC - 1] = w2 >> 16;
            } else {
                gw[0] = a.geo[i];
            }
            const V<VEC> vol = ldv<VEC>(a.vol, i), Qo = ldv<VEC>(a.Q_in, i);
            V<VEC> wid, tn, cs, rg, sn;
            if (GP) {
                wid = ldv<VEC>(a.width.ptr, i); tn = ldv<VEC>(a.tan_angle.ptr, i);
                cs = ldv<VEC>(a.cos_angle.ptr, i); rg = ldv<VEC>(a.rough.ptr, i);
#pragma unroll
                for (int j = 0; j < VEC; ++j) sn.v[j] = a.sinu.val;
            } else {
                wid = ldsg<VEC>(a.width, i); tn = ldsg<VEC>(a.tan_angle, i);
                cs = ldsg<VEC>(a.cos_angle, i); rg = ldsg<VEC>(a.rough, i);
                sn = ldsg<VEC>(a.sinu, i);
            }
            const V<VEC> sl = ldv<VEC>((DW || DIAG || LOW) ? a.S_bed : a.sqrtS, i);
            V<VEC> Iin, hin, Iout, hout;
            const bool f_infil = SRC && (a.flags & TFB_CH_SRC_INFIL);
            const bool f_snow = SRC && (a.flags & TFB_CH_SRC_SNOW);
            if (f_infil) Iin = ldv<VEC>(a.I, i);
            if (f_snow) hin = ldv<VEC>(a.h_swe, i);
            const bool f_flood = (a.flags & TFB_CH_FLOOD) != 0u;
            V<VEC> vbk, o_df, o_vf;
            if (f_flood) vbk = ldsg<VEC>(a.vol_bankfull, i);
            const float dxr = __ldg(a.dx + r), dyr = __ldg(a.dy + r), ddr = __ldg(a.dd + r);
            const double dar = (GP || !a.da.ptr) ? a.da.val : __ldg(a.da.ptr + r);
            // ---- children's discharge: predicated loads --------------------
            double Qc[VEC][8];
#pragma unroll
            for (int j = 0; j < VEC; ++j)
                load_children(a.Q_in + i + j, a.pitch, gw[j] >> 8, Qc[j]);
            V<VEC> o_vol, o_d, o_u, o_Q;
            V<VEC> o_Rh, o_A, o_P, o_f, o_tau, o_us, o_fr, o_S, o_R;
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                Cell g;
                g.code = gw[j] & 0xffu;
                g.cmask = gw[j] >> 8;
                g.ds = sn.v[j] * (double)ds_of(g.code, dxr, dyr, ddr);
                g.da = dar;
                g.width = wid.v[j]; g.tan_a = tn.v[j]; g.cos_a = cs.v[j];
                g.vol = vol.v[j]; g.Q = Qo.v[j];
                g.I = f_infil ? Iin.v[j] : 0.0;
                g.h_swe = f_snow ? hin.v[j] : 0.0;
                g.vb = f_flood ? vbk.v[j] : 0.0;
                double v, d, df, R;
                SrcState ns;
                vol_depth<SRC, UR, true>(a, k, i + j, g, Qc[j], v, d, df, R, ns, &acc);
                if (SRC) { Iout.v[j] = ns.I; hout.v[j] = ns.h_swe; }
                CellOut o;
                finish_cell<DW, LOW, DIAG, SRC, UR>(a, k, r, c0 + j, i + j, g, v, d, df, sl.v[j],
                                                rg.v[j], acc, o);
                o_df.v[j] = o.df; o_vf.v[j] = o.vf;
                o_vol.v[j] = o.vol; o_d.v[j] = o.d; o_u.v[j] = o.u; o_Q.v[j] = o.Qn;
                if (DIAG) {
                    o_Rh.v[j] = o.Rh; o_A.v[j] = o.A_wet; o_P.v[j] = o.P_wet; o_f.v[j] = o.f;
                    o_tau.v[j] = o.tau; o_us.v[j] = o.u_star; o_fr.v[j] = o.froude;
                    o_S.v[j] = o.S;
                }
                o_R.v[j] = R;
            }
            if (f_infil) stv<VEC>(a.I_out, i, Iout);
            if (f_snow) {
                stv<VEC>(a.h_swe_out, i, hout);
                if (a.h_snow) {
                    V<VEC> hs;
#pragma unroll
                    for (int j = 0; j < VEC; ++j) hs.v[j] = hout.v[j] * (a.rho_H2O / a.rho_snow);
                    stv<VEC>(a.h_snow, i, hs);
                }
            }
            if (f_flood) {
                if (a.d_flood) stv<VEC>(a.d_flood, i, o_df);
                if (a.vol_flood) stv<VEC>(a.vol_flood, i, o_vf);
            }
            stv<VEC>(a.vol_out, i, o_vol);
            stv<VEC>(a.d, i, o_d);
            stv<VEC>(a.u, i, o_u);
            stv<VEC>(a.Q_out, i, o_Q);
            if (DIAG) {
                stv<VEC>(a.Rh, i, o_Rh); stv<VEC>(a.A_wet, i, o_A); stv<VEC>(a.P_wet, i, o_P);
                stv<VEC>(a.f, i, o_f); stv<VEC>(a.tau, i, o_tau); stv<VEC>(a.u_star, i, o_us);
                stv<VEC>(a.froude, i, o_fr);
                if (DW && a.S_free) stv<VEC>(a.S_free, i, o_S);
            }
            if ((a.flags & TFB_CH_WRITE_R) && a.R) stv<VEC>(a.R, i, o_R);
        }
    }

    // ---- CTA reduction: warp shuffles only where a warp has something -------
    __shared__ double sm[kBY][kPartialStride];
    __shared__ bool is_last;
    const int lane = threadIdx.x, warp = threadIdx.y;
#pragma unroll
    for (int j = 0; j < kNSum; ++j) {
        double v = acc.s[j];
        if (__any_sync(0xffffffffu, v != 0.0)) v = warp_sum(v);
        if (lane == 0) sm[warp][j] = v;
    }
    {
        unsigned any = 0;
#pragma unroll
        for (int j = 0; j < kNCnt; ++j) any |= acc.n[j];
        const bool some = __any_sync(0xffffffffu, any != 0u);
#pragma unroll
        for (int j = 0; j < kNCnt; ++j) {
            unsigned v = acc.n[j];
            if (some) v = __reduce_add_sync(0xffffffffu, v);
            if (lane == 0) sm[warp][kNSum + j] = (double)v;
        }
        const double m = warp_min((double)acc.dt_min);
        if (lane == 0) sm[warp][P_DTMIN] = m;
    }
    __syncthreads();
    const int tid = warp * kBX + lane;
    const int nblk = gridDim.x;
    if (tid < kPartialStride) {
        double v = 0.0;
        if (tid < S_NSLOT) {
            for (int w = 0; w < kBY; ++w) v += sm[w][tid];
        } else if (tid == P_DTMIN) {
            v = sm[0][tid];
            for (int w = 1; w < kBY; ++w) v = fmin(v, sm[w][tid]);
        }
        a.partials[(int64_t)(blockIdx.x + 2) * kPartialStride + tid] = v;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned tk = atomicAdd(a.ticket, 1u);
        is_last = (tk == (unsigned)nblk - 1u);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // last CTA: thread j folds slot j over all CTAs in index order (deterministic)
    __shared__ double tot[kPartialStride];
    if (tid < kPartialStride) {
        const volatile double *P = a.partials;
        double v = (tid == P_DTMIN) ? 3.0e38 : 0.0;
        if (tid < S_NSLOT || tid == P_DTMIN) {
            for (int b = 0; b < nblk; ++b) {
                const double x = P[(int64_t)(b + 2) * kPartialStride + tid];
                v = (tid == P_DTMIN) ? fmin(v, x) : (v + x);
            }
        }
        tot[tid] = v;
    }
    __syncthreads();
    if (tid == 0) {
        double *outp = a.partials + kPartialStride;       // row 1
        const bool has_out = (outp[P_HASOUT] != 0.0);
        double *h = a.partials;                           // row 0: this rank's sums
        for (int j = 0; j < S_NSLOT; ++j) h[j] = tot[j];
        h[P_HASOUT_SUM] = has_out ? 1.0 : 0.0;
        h[P_DTMIN] = tot[P_DTMIN];
        h[P_QOUT] = has_out ? outp[P_QOUT] : 0.0;
        h[P_UOUT] = has_out ? outp[P_UOUT] : 0.0;
        h[P_DOUT] = has_out ? outp[P_DOUT] : 0.0;
        h[P_FOUT] = has_out ? outp[P_FOUT] : 0.0;
        h[P_HASOUT] = 0.0; h[P_FAIL] = 0.0; h[23] = 0.0;
        outp[P_HASOUT] = 0.0;
        if (a.finalize) fold_scalars(a, h);
        *a.ticket = 0u;    // re-arm for the next step
    }
}

__global__ void channels_fold_kernel(const tfb_channels_step_args a, const double *sums) {
    if (threadIdx.x == 0 && blockIdx.x == 0) fold_scalars(a, sums);
}

static int persistent_grid(int64_t n_tiles, int ctas_per_sm) {
    int64_t g = (int64_t)sm_count() * ctas_per_sm;
    if (g > n_tiles) g = n_tiles;
    return (int)(g < 1 ? 1 : g);
}

template <int VEC, bool DW, bool LOW, bool DIAG, bool SRC, bool GP, bool UR>
static int launch(const tfb_channels_step_args &a, const ChConsts &k, cudaStream_t st) {
    const int tw = kBX * VEC;
    const int64_t n_tiles = (int64_t)((a.nx + tw - 1) / tw) * ((a.ny + kTileRows - 1) / kTileRows);
    int per_sm = 0;
    TFB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
        &per_sm, channels_step_kernel<VEC, DW, LOW, DIAG, SRC, GP, UR>, kBX * kBY, 0));
    if (per_sm < 1) per_sm = 1;
    dim3 block(kBX, kBY);
    channels_step_kernel<VEC, DW, LOW, DIAG, SRC, GP, UR>
        <<<persistent_grid(n_tiles, per_sm), block, 0, st>>>(a, k);
    TFB_CUDA_CHECK(cudaGetLastError());
    return TFB_OK;
}

template <int VEC>
static int dispatch(const tfb_channels_step_args &a, ChConsts &k, cudaStream_t st) {
    const bool dw = a.flags & TFB_CH_DIFFUSIVE, low = a.flags & TFB_CH_LAW_OF_WALL;
    const bool diag = a.flags & TFB_CH_DIAG;
    const bool src = a.flags & (TFB_CH_SRC_MET | TFB_CH_SRC_SNOW | TFB_CH_SRC_EVAP |
                                TFB_CH_SRC_INFIL);
    // fast-path eligibility
    const bool gp = (VEC == 2) && !low && a.width.ptr && a.tan_angle.ptr && a.cos_angle.ptr &&
                    a.rough.ptr && !a.sinu.ptr && !a.da.ptr;
    const bool ur = gp && !src && !a.P_rain.ptr && !a.SM.ptr && !a.GW.ptr && !a.MR.ptr &&
                    !a.IN.ptr && !a.ET.ptr && !a.is_R_grid && !(a.flags & TFB_CH_WRITE_R);
    if (ur) {
        const double R = (((a.P_rain.val + a.SM.val) + a.GW.val) + a.MR.val) - (0.0 + a.IN.val);
        k.R_uniform = R;
        k.Rdadt_uniform = (R * a.da.val) * a.dt;
    }
#define TFB_GENERIC(DW_, LOW_, DIAG_, SRC_)                                        \
    if (dw == DW_ && low == LOW_ && diag == DIAG_ && src == SRC_)                  \
        return launch<VEC, DW_, LOW_, DIAG_, SRC_, false, false>(a, k, st);
#define TFB_FAST(DW_, DIAG_, SRC_, UR_)                                            \
    if (VEC == 2 && gp && ur == UR_ && dw == DW_ && diag == DIAG_ && src == SRC_)  \
        return launch<2, DW_, false, DIAG_, SRC_, true, UR_>(a, k, st);
    TFB_FAST(false, false, false, true)  TFB_FAST(false, false, false, false)
    TFB_FAST(false, false, true, false)  TFB_FAST(false, true, false, true)
    TFB_FAST(false, true, false, false)  TFB_FAST(false, true, true, false)
    TFB_FAST(true, false, false, true)   TFB_FAST(true, false, false, false)
    TFB_FAST(true, false, true, false)   TFB_FAST(true, true, false, true)
    TFB_FAST(true, true, false, false)   TFB_FAST(true, true, true, false)
    TFB_GENERIC(false, false, false, false) TFB_GENERIC(false, false, false, true)
    TFB_GENERIC(false, false, true, false)  TFB_GENERIC(false, false, true, true)
    TFB_GENERIC(false, true, false, false)  TFB_GENERIC(false, true, false, true)
    TFB_GENERIC(false, true, true, false)   TFB_GENERIC(false, true, true, true)
    TFB_GENERIC(true, false, false, false)  TFB_GENERIC(true, false, false, true)
    TFB_GENERIC(true, false, true, false)   TFB_GENERIC(true, false, true, true)
    TFB_GENERIC(true, true, false, false)   TFB_GENERIC(true, true, false, true)
    TFB_GENERIC(true, true, true, false)    TFB_GENERIC(true, true, true, true)
#undef TFB_GENERIC
#undef TFB_FAST
    return TFB_ERR_ARG;
}

}  // namespace tfb

using namespace tfb;

extern "C" int64_t tfb_channels_partials_len(int32_t ny, int32_t nx) {
    (void)ny; (void)nx;
    // one row per resident CTA (<= 8 per SM) + outlet row + rank-sum row
    return ((int64_t)sm_count() * 8 + 2) * kPartialStride;
}

extern "C" int tfb_channels_pack_geo(const uint8_t *code, int32_t ny, int32_t nx,
                                     int32_t halo, uint16_t *geo, void *stream) {
    TFB_ARG_CHECK(code && geo && ny > 0 && nx > 0 && halo >= 0);
    dim3 b(32, 8), g((nx + 31) / 32, (ny + 2 * halo + 7) / 8);
    pack_geo_kernel<<<g, b, 0, (cudaStream_t)stream>>>(code, ny, nx, halo, geo);
    TFB_CUDA_CHECK(cudaGetLastError());
    return TFB_OK;
}

extern "C" int tfb_channels_step(const tfb_channels_step_args *ap, void *stream) {
    TFB_ARG_CHECK(ap != nullptr);
    tfb_channels_step_args a = *ap;
    if (a.pitch == 0) a.pitch = a.nx;                  // dense rows
    TFB_ARG_CHECK(a.ny > 0 && a.nx > 0 && a.halo >= 0 && a.pitch >= a.nx);
    TFB_ARG_CHECK(a.geo && a.dx && a.dy && a.dd);
    TFB_ARG_CHECK(a.vol && a.vol_out && a.Q_in && a.Q_out && a.d && a.u);
    TFB_ARG_CHECK(a.Q_in != a.Q_out);
    TFB_ARG_CHECK(a.scalars && a.partials && a.ticket);
    TFB_ARG_CHECK(a.dt > 0.0);
    const bool dw = a.flags & TFB_CH_DIFFUSIVE, diag = a.flags & TFB_CH_DIAG;
    const bool src = a.flags & (TFB_CH_SRC_MET | TFB_CH_SRC_SNOW | TFB_CH_SRC_EVAP |
                                TFB_CH_SRC_INFIL);
    if (dw || diag || (a.flags & TFB_CH_LAW_OF_WALL)) TFB_ARG_CHECK(a.S_bed != nullptr);
    else TFB_ARG_CHECK(a.sqrtS != nullptr);
    if (diag)
        TFB_ARG_CHECK(a.Rh && a.A_wet && a.P_wet && a.f && a.tau && a.u_star && a.froude);
    if (a.flags & TFB_CH_SRC_SNOW) TFB_ARG_CHECK(a.h_swe && a.h_swe_out);
    if (a.flags & TFB_CH_SRC_INFIL) TFB_ARG_CHECK(a.I && a.I_out);
    if (dw) {
        // neighbours re-read pre-step state: it must not be updated in place
        TFB_ARG_CHECK(a.vol != a.vol_out);
        if (a.flags & TFB_CH_SRC_SNOW) TFB_ARG_CHECK(a.h_swe != a.h_swe_out);
        if (a.flags & TFB_CH_SRC_INFIL) TFB_ARG_CHECK(a.I != a.I_out);
    }
    if (src) TFB_ARG_CHECK(a.is_R_grid == 1);
    if (a.flags & TFB_CH_FLOOD)
        TFB_ARG_CHECK(a.flood_manning_n > 0.0 && a.width_ratio > 0.0);
    if (a.flags & TFB_CH_ATTENUATE) {
        TFB_ARG_CHECK(a.vol_stored && a.vol_stored_out && a.t_drain > 0.0);
        if (dw) TFB_ARG_CHECK(a.vol_stored != a.vol_stored_out);
    }
    ChConsts k;
    k.g = 9.81; k.aval = 0.476; k.kappa = 0.408;
    k.law_const = sqrt(k.g) / k.kappa;                 // channels_base :531
    cudaStream_t st = (cudaStream_t)stream;
    // 128-bit path needs every row 16-byte aligned: even nx and aligned bases
    auto al16 = [](const void *p) { return p == nullptr || (((uintptr_t)p) & 15u) == 0; };
    const bool vec2 = (a.nx % 2 == 0) && (a.pitch % 2 == 0) && al16(a.vol) && al16(a.vol_out) && al16(a.Q_in) &&
                      al16(a.Q_out) && al16(a.d) && al16(a.u) && al16(a.S_bed) &&
                      al16(a.sqrtS) && al16(a.width.ptr) && al16(a.tan_angle.ptr) &&
                      al16(a.cos_angle.ptr) && al16(a.rough.ptr) && al16(a.sinu.ptr) &&
                      al16(a.Rh) && al16(a.A_wet) && al16(a.P_wet) && al16(a.f) &&
                      al16(a.tau) && al16(a.u_star) && al16(a.froude) && al16(a.S_free) &&
                      al16(a.R) && al16(a.I) && al16(a.I_out) && al16(a.h_swe) && al16(a.h_swe_out) &&
                      al16(a.h_snow) && al16(a.vol_bankfull.ptr) && al16(a.d_flood) && al16(a.vol_flood) &&
                      ((((uintptr_t)a.geo) & 3u) == 0);
    return vec2 ? dispatch<2>(a, k, st) : dispatch<1>(a, k, st);
}

extern "C" int tfb_channels_fold(const tfb_channels_step_args *ap, const double *sums,
                                 void *stream) {
    TFB_ARG_CHECK(ap != nullptr && sums != nullptr && ap->scalars != nullptr);
    channels_fold_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(*ap, sums);
    TFB_CUDA_CHECK(cudaGetLastError());
    return TFB_OK;
}
