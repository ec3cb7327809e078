// Stand-alone source-term kernels: one per reference component update().
// Used when the source components run as separate BMI components (EMELI-style
// coupling, each at its own dt); the routing kernel can alternatively evaluate
// the same expressions fused (TFB_CH_SRC_* flags in tfb_channels.cu).
//
// Each kernel is a single coalesced pass over the grid with a deterministic
// block-partial -> last-block reduction for the volume integrals.
#include "tfb_common.cuh"

#include <math.h>

namespace tfb {

constexpr int kThreads = 256;

// grid-stride element loop + deterministic reduction of up to 4 sums
template <int NS, typename F>
__device__ __forceinline__ void reduce_and_fold(double (&acc)[NS], double *partials,
                                                unsigned *ticket, double *sums_out, F) {
    __shared__ double sm[kThreads / 32][NS];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < NS; ++j) {
        const double v = warp_sum(acc[j]);
        if (lane == 0) sm[warp][j] = v;
    }
    __syncthreads();
    if (threadIdx.x < NS) {
        double v = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) v += sm[w][threadIdx.x];
        partials[(int64_t)blockIdx.x * NS + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1u);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    __shared__ double red[kThreads];
    const volatile double *P = partials;
    for (int j = 0; j < NS; ++j) {
        double v = 0.0;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += kThreads) v += P[(int64_t)b * NS + j];
        red[threadIdx.x] = v;
        __syncthreads();
        for (int s = kThreads / 2; s > 0; s >>= 1) {
            if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
            __syncthreads();
        }
        if (threadIdx.x == 0 && sums_out) sums_out[j] += red[0];
        __syncthreads();
    }
    if (threadIdx.x == 0) *ticket = 0u;
}

__device__ __forceinline__ double da_at(const tfb_sg &da, int64_t i, int nx) {
    return da.ptr ? __ldg(da.ptr + (i / nx)) : da.val;
}

// ---- met: P -> P_rain, P_snow; vol_P, vol_PR, vol_PS (met_base.py:1247-1446)
__global__ void __launch_bounds__(kThreads)
met_precip_kernel(tfb_sg P, tfb_sg T_air, double T_rs, tfb_sg da, double dt, int64_t n,
                  int nx, double *P_rain, double *P_snow, double *sums, double *partials,
                  unsigned *ticket) {
    double acc[3] = {0.0, 0.0, 0.0};
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * kThreads) {
        const double p = sg_at(P, i), ta = sg_at(T_air, i), a = da_at(da, i, nx);
        const double pr = p * ((ta > T_rs) ? 1.0 : 0.0);
        const double ps = p * ((ta <= T_rs) ? 1.0 : 0.0);
        if (P_rain) P_rain[i] = pr;
        if (P_snow) P_snow[i] = ps;
        acc[0] += (p * a) * dt; acc[1] += (pr * a) * dt; acc[2] += (ps * a) * dt;
    }
    reduce_and_fold<3>(acc, partials, ticket, sums, 0);
}

// ---- infil: Green-Ampt v1 (infil_green_ampt.py:511-578; infil_base.py:562-811)
__global__ void __launch_bounds__(kThreads)
green_ampt_kernel(tfb_sg P_rain, tfb_sg SM, tfb_sg Ks, tfb_sg Ki, tfb_sg qs, tfb_sg qi,
                  tfb_sg G, tfb_sg h_table, tfb_sg elev, int use_table, tfb_sg da,
                  double dt, int64_t n, int nx, double *IN, double *I, double *sums,
                  double *partials, unsigned *ticket) {
    double acc[2] = {0.0, 0.0};
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * kThreads) {
        const double ks = sg_at(Ks, i);
        const double P_total = sg_at(P_rain, i) + sg_at(SM, i);     // :571
        const double dK = ks - sg_at(Ki, i);
        const double dq = sg_at(qs, i) - sg_at(qi, i);
        const double Iv = I[i];
        const double fc = (((sg_at(G, i) * dK) * dq) / Iv) + ks;    // :540-542
        double in = np_min(fc, P_total);                            // :567
        if (P_total < ks) in = P_total;                             // :984-993
        if (use_table && sg_at(h_table, i) >= sg_at(elev, i)) in = 0.0;   // :697-701
        IN[i] = in;
        I[i] = Iv + (in * dt);                                      // :797
        acc[0] += (in * da_at(da, i, nx)) * dt;                     // :732-736
        if (!isfinite(in)) acc[1] += 1.0;                           // :906-950
    }
    reduce_and_fold<2>(acc, partials, ticket, sums, 0);
}

// ---- snow: degree-day (snow_degree_day.py:293-360; snow_base.py:490-703)
__global__ void __launch_bounds__(kThreads)
degree_day_kernel(tfb_sg P_snow, tfb_sg T_air, tfb_sg c0, tfb_sg T0, double ratio,
                  tfb_sg da, double dt, int64_t n, int nx, double *SM, double *h_swe,
                  double *h_snow, double *sums, double *partials, unsigned *ticket) {
    double acc[1] = {0.0};
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * kThreads) {
        const double h = h_swe[i];
        double M = (sg_at(c0, i) / 8.64E7) * (sg_at(T_air, i) - sg_at(T0, i));   // :339
        M = np_min(M, h / dt);                                      // :490-512
        M = np_max(M, 0.0);
        SM[i] = M;
        acc[0] += (M * da_at(da, i, nx)) * dt;
        double hn = h + (sg_at(P_snow, i) * dt);                    // :559-609
        hn = hn - (M * dt);
        hn = np_max(hn, 0.0);
        h_swe[i] = hn;
        if (h_snow) h_snow[i] = hn * ratio;                         // :631-703
    }
    reduce_and_fold<1>(acc, partials, ticket, sums, 0);
}

// ---- evap: Priestley-Taylor (evap_priestley_taylor.py:308-413; evap_base :326-377)
__global__ void __launch_bounds__(kThreads)
priestley_taylor_kernel(tfb_sg T_air, tfb_sg T_surf, tfb_sg Qn_SW, tfb_sg Qn_LW,
                        double alpha, double K_soil, double soil_x, double T_soil_x,
                        tfb_sg da, double dt, int64_t n, int nx, double *Qc_out,
                        double *ET, double *sums, double *partials, unsigned *ticket) {
    double acc[1] = {0.0};
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * kThreads) {
        const double Qc = K_soil * (sg_at(T_surf, i) - T_soil_x) / soil_x;
        const double Qnet = sg_at(Qn_SW, i) + sg_at(Qn_LW, i);
        const double Qet = (alpha * (0.406 + (0.011 * sg_at(T_air, i)))) * (Qnet + Qc);
        const double et = np_max(Qet / 2.5E+9, 0.0);
        if (Qc_out) Qc_out[i] = Qc;
        ET[i] = et;
        acc[0] += (et * da_at(da, i, nx)) * dt;
    }
    reduce_and_fold<1>(acc, partials, ticket, sums, 0);
}

// ---- channels finalize: interior min/max of Q,u,d + total volume
// (channels_base.update_mins_and_maxes :2986-3081 uses [1:-1, 1:-1])
__global__ void __launch_bounds__(kThreads)
minmax_kernel(const double *Q, const double *u, const double *d, const double *vol, int ny,
              int nx, double *out6, double *vol_sum, double *partials, unsigned *ticket) {
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    double acc = 0.0;
    const int64_t n = (int64_t)ny * nx;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * kThreads) {
        acc += vol[i];
        const int r = (int)(i / nx), c = (int)(i - (int64_t)r * nx);
        if (r == 0 || r == ny - 1 || c == 0 || c == nx - 1) continue;
        const double v[3] = {Q[i], u[i], d[i]};
#pragma unroll
        for (int j = 0; j < 3; ++j) { mn[j] = fmin(mn[j], v[j]); mx[j] = fmax(mx[j], v[j]); }
    }
    __shared__ double sm[kThreads / 32][7];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double a = warp_min(mn[j]), b = warp_max(mx[j]);
        if (lane == 0) { sm[warp][2 * j] = a; sm[warp][2 * j + 1] = b; }
    }
    { const double s = warp_sum(acc); if (lane == 0) sm[warp][6] = s; }
    __syncthreads();
    if (threadIdx.x < 7) {
        double v = sm[0][threadIdx.x];
        for (int w = 1; w < kThreads / 32; ++w) {
            const double x = sm[w][threadIdx.x];
            v = (threadIdx.x == 6) ? v + x : ((threadIdx.x & 1) ? fmax(v, x) : fmin(v, x));
        }
        partials[(int64_t)blockIdx.x * 7 + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1u);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x < 7) {
        const volatile double *P = partials;
        double v = P[threadIdx.x];
        for (int b = 1; b < (int)gridDim.x; ++b) {
            const double x = P[(int64_t)b * 7 + threadIdx.x];
            v = (threadIdx.x == 6) ? v + x : ((threadIdx.x & 1) ? fmax(v, x) : fmin(v, x));
        }
        if (threadIdx.x == 6) *vol_sum = v; else out6[threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) *ticket = 0u;
}

// ---- met: energy chain (met_base.py update :790-809; solar_funcs.py) -----------
struct MetDay { double delta, E0, sin_d, cos_d, tan_d; };

constexpr double kPi = 3.141592653589793;
constexpr double kOmegaH = (360.0 / 24.0) * (kPi / 180.0);     // solar_funcs.py:236-246
constexpr double kIsc = 1361.5;                                // :133

__device__ __forceinline__ double sun_offset(double lat_deg, double tan_d, double sign) {
    const double lat_rad = lat_deg * (kPi / 180.0);            // :286-340
    double arg = (-1.0 * tan(lat_rad)) * tan_d;
    arg = np_min(np_max(-1.0, arg), 1.0);
    return (sign * acos(arg)) / kOmegaH;
}

__global__ void __launch_bounds__(kThreads)
met_energy_kernel(const tfb_met_energy_args a, const MetDay dy) {
    const int64_t n = (int64_t)a.ny * a.nx;
    const double g = 9.81, kappa = 0.408, rho_air = 1.2614, Cp_air = 1005.7, Lv = 2500000.0;
    const double sigma = 5.67E-8, C_to_K = 273.15;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * kThreads) {
        const double Ta = sg_at(a.T_air, i), RH = sg_at(a.RH, i), uz = sg_at(a.uz, i);
        const double z = sg_at(a.z, i), z0 = sg_at(a.z0_air, i), hs = sg_at(a.h_snow, i);
        const double hi = sg_at(a.h_ice, i), p0 = sg_at(a.p0, i), alb = sg_at(a.albedo, i);
        const double ems = sg_at(a.em_surf, i), lat = sg_at(a.lat_deg, i);
        const double al = sg_at(a.alpha, i), be = sg_at(a.beta, i), th = sg_at(a.TSN_offset, i);
        const double Cf = sg_at(a.cloud_factor, i), Ff = sg_at(a.canopy_factor, i);
        // saturation vapour pressure [mbar] :1657-1741
        auto e_sat = [&](double T) {
            double e;
            if (!a.satterlund) {
                const double term1 = (17.3 * T) / (T + 237.3);
                e = 0.611 * exp(term1);
            } else {
                const double term1 = 2353.0 / (T + 273.15);
                e = pow(10.0, 11.4 - term1);
                e = e / 1000.0;
            }
            return e * 10.0;
        };
        const double esa = e_sat(Ta);
        const double ea = RH * esa;                                  // :1743-1788
        const double lt = log(ea / 6.1121);                          // :1790-1835
        const double Td = (257.14 * lt) / (18.678 - lt);
        const double Ts = (hs > 0.0 || hi > 0.0) ? np_min(Td, 0.0) : Td;   // :1837-1853
        const double ess = e_sat(Ts);
        const double top = (g * z) * (Ta - Ts);                      // :1448-1475
        const double bot = (uz * uz) * (Ta + 273.15);
        const double Ri = top / bot;
        const double arg = kappa / log((z - hs) / z0);               // :1477-1625
        const double Dn = uz * (arg * arg);
        const double Dh = (Ri > 0.0) ? (Dn / (1.0 + (10.0 * Ri))) : (Dn * (1.0 - (10.0 * Ri)));
        const double Qh = ((rho_air * Cp_air) * Dh) * (Ta - Ts);     // :1627-1655
        const double Wp = 1.12 * exp(0.0614 * Td);                   // :1855-1873
        const double es = RH * ess;
        const double Qe = (((rho_air * Lv) * Dh) * (ea - es)) * (0.622 / p0);   // :1875-1936
        // ---- clear-sky radiation on the sloping cell, solar_funcs.py:870-942 ------------
        const double lat_rad = lat * (kPi / 180.0);
        const double sl = sin(lat_rad), cl = cos(lat_rad);
        const double coth = cos(kOmegaH * th);
        const double Z = acos((sl * dy.sin_d) + ((cl * dy.cos_d) * coth));      // :248-269
        const double gam = np_max(90.0 - (Z * (180.0 / kPi)), 0.0);            // :477-551
        const double M_opt = 1.0 / (sin(gam * (kPi / 180.0)) + (0.50572 / pow(gam + 6.07995, 1.6364)));
        const double a_sa = -0.1240 - (0.0207 * Wp), b_sa = -0.0682 - (0.0248 * Wp);   // :567-594
        const double tau = np_min(np_max(exp(a_sa + (b_sa * M_opt)) - a.dust_atten, 0.0), 1.0);
        const double sb = sin(be), cb = cos(be), sa = sin(al), ca = cos(al);
        const double lat_eq = asin(((sb * ca) * cl) + (cb * sl));               // :723-751
        const double dlon = atan((sb * sa) / ((cb * cl) - ((sb * sl) * ca)));   // :701-721
        const double t1 = dy.cos_d * cos(lat_eq);                               // :822-868
        const double t2 = cos((kOmegaH * th) + dlon);
        const double t3 = sin(lat_eq) * dy.sin_d;
        const double K_ETs = np_max((kIsc * dy.E0) * ((t1 * t2) + t3), 0.0);
        const double a_s = -0.0363 - (0.0084 * Wp), b_s = -0.0572 - (0.0173 * Wp);   // :619-636
        const double gam_s = (1.0 - exp(a_s + (b_s * M_opt))) + a.dust_atten;
        const double K_ET = np_max((kIsc * dy.E0) * (((dy.cos_d * cl) * coth) + (dy.sin_d * sl)), 0.0);
        const double K_dif = (0.5 * gam_s) * K_ET;                              // :638-651
        const double K_glob = (tau * K_ET) + K_dif;                             // :653-669
        const double K_bs = ((0.5 * gam_s) * alb) * K_glob;                     // :671-699
        double K_cs = ((tau * K_ETs) + K_dif) + K_bs;
        const double eq_deg = lat_eq * (180.0 / kPi);
        const double t_noon = (-1.0 * dlon) / kOmegaH;                          // :753-761
        const double T_sr = np_max(sun_offset(eq_deg, dy.tan_d, -1.0) + t_noon,
                                   sun_offset(lat, dy.tan_d, -1.0));            // :763-786
        const double T_ss = np_min(sun_offset(eq_deg, dy.tan_d, 1.0) + t_noon,
                                   sun_offset(lat, dy.tan_d, 1.0));             // :788-811
        if (th <= T_sr || th >= T_ss) K_cs = 0.0;
        const double Qsw = K_cs * (1.0 - alb);                                  // met_base :2048-2088
        // air emissivity and net longwave :2090-2224
        const double TaK = Ta + C_to_K, TsK = Ts + C_to_K;
        double em;
        if (!a.satterlund) {
            const double term1 = ((1.0 - Ff) * 1.72) * pow((ea / 10.0) / TaK, 1.0 / 7.0);
            const double term2 = 1.0 + (0.22 * (Cf * Cf));
            em = (term1 * term2) + Ff;
        } else {
            em = 1.08 * (1.0 - exp(-1.0 * pow(ea, TaK / 2016.0)));
        }
        const double LW_in = (em * sigma) * pow(TaK, 4.0);
        double LW_out = (ems * sigma) * pow(TsK, 4.0);
        LW_out = LW_out + ((1.0 - ems) * LW_in);
        const double Qlw = LW_in - LW_out;
        const double Qsum = ((((Qsw + Qlw) + Qh) + Qe) + sg_at(a.Qa, i)) + sg_at(a.Qc, i);   // :2305
        if (a.e_sat_air) a.e_sat_air[i] = esa;
        if (a.e_air) a.e_air[i] = ea;
        if (a.T_dew) a.T_dew[i] = Td;
        if (a.T_surf) a.T_surf[i] = Ts;
        if (a.e_sat_surf) a.e_sat_surf[i] = ess;
        if (a.Ri) a.Ri[i] = Ri;
        if (a.Dn) a.Dn[i] = Dn;
        if (a.Dh) a.Dh[i] = Dh;
        if (a.Qh) a.Qh[i] = Qh;
        if (a.W_p) a.W_p[i] = Wp;
        if (a.e_surf) a.e_surf[i] = es;
        if (a.Qe) a.Qe[i] = Qe;
        if (a.Qn_SW) a.Qn_SW[i] = Qsw;
        if (a.em_air) a.em_air[i] = em;
        if (a.Qn_LW) a.Qn_LW[i] = Qlw;
        if (a.Q_sum) a.Q_sum[i] = Qsum;
    }
}

static int n_blocks(int64_t n) {
    int64_t want = (n + kThreads - 1) / kThreads;
    int64_t cap = (int64_t)sm_count() * 8;      // 8 resident CTAs of 256 per SM
    return (int)(want < cap ? want : cap);
}

// ---- albedo (met_base.update_albedo :1995-2045) -----------------------------------------
// 'aging' (Rohrer & Braun 1994): albedo = 0.4 + 0.44 exp(-n r) where snow lies, n = days
// since the snowfall of the last 3 days (a ring of `slots` per-step snow depths) reached
// 3 cm; 0.3 on bare ice, 0.15 on bare ground.  The reference rolls a (slots, ny, nx) array
// and sums it over time with np.sum: a sequential sum oldest -> newest, kept here by
// walking the ring from the slot after `head`.
__global__ void __launch_bounds__(256)
met_albedo_kernel(tfb_sg P_snow, tfb_sg T_air, tfb_sg h_snow, tfb_sg h_ice, int64_t ncell,
                  double *__restrict__ ring, int slots, int head, double *__restrict__ n_days,
                  double dt, double ws_ratio, double days_per_dt, int aging,
                  double *__restrict__ albedo) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncell) return;
    const double hs = sg_at(h_snow, i), hi = sg_at(h_ice, i);
    double a = albedo[i];
    if (aging) {
        ring[(int64_t)head * ncell + i] = (sg_at(P_snow, i) * dt) * ws_ratio;       // newest (:2017)
        int s = head + 1;
        if (s == slots) s = 0;
        double total = ring[(int64_t)s * ncell + i];                              // oldest
        for (int k = 1; k < slots; ++k) {
            if (++s == slots) s = 0;
            total += ring[(int64_t)s * ncell + i];
        }
        double n = n_days[i];
        if (total >= 0.03) n = 0.0;
        else if (total < 0.03) n = n + days_per_dt;
        n_days[i] = n;
        const double r = (sg_at(T_air, i) > 0.0) ? 0.12 : 0.05;
        const double snow_albedo = 0.4 + 0.44 * exp(-n * r);
        if (hs > 0.0) a = snow_albedo;
    } else {
        if (hs > 0.0) a = 0.75;
    }
    if (hs == 0.0 && hi > 0.0) a = 0.3;
    if (hs == 0.0 && hi == 0.0) a = 0.15;
    albedo[i] = a;
}

}  // namespace tfb

using namespace tfb;

extern "C" int64_t tfb_sources_partials_len(void) {
    return (int64_t)sm_count() * 8 * 8;
}

extern "C" int tfb_met_precip(tfb_sg P, tfb_sg T_air, double T_rain_snow, tfb_sg da,
                              double dt, int32_t ny, int32_t nx, double *P_rain,
                              double *P_snow, double *sums3, double *partials,
                              uint32_t *ticket, void *stream) {
    TFB_ARG_CHECK(ny > 0 && nx > 0 && partials && ticket);
    const int64_t n = (int64_t)ny * nx;
    met_precip_kernel<<<n_blocks(n), kThreads, 0, (cudaStream_t)stream>>>(
        P, T_air, T_rain_snow, da, dt, n, nx, P_rain, P_snow, sums3, partials, ticket);
    TFB_CUDA_CHECK(cudaGetLastError());
    return TFB_OK;
}

extern "C" int tfb_infil_green_ampt(tfb_sg P_rain, tfb_sg SM, tfb_sg Ks, tfb_sg Ki,
                                    tfb_sg qs, tfb_sg qi, tfb_sg G, tfb_sg h_table,
                                    tfb_sg elev, int32_t use_table, tfb_sg da, double dt,
                                    int32_t ny, int32_t nx, double *IN, double *I,
                                    double *sums2, double *partials, uint32_t *ticket,
                                    void *stream) {
    TFB_ARG_CHECK(ny > 0 && nx > 0 && IN && I && partials && ticket);
    const int64_t n = (int64_t)ny * nx;
    green_ampt_kernel<<<n_blocks(n), kThreads, 0, (cudaStream_t)stream>>>(
        P_rain, SM, Ks, Ki, qs, qi, G, h_table, elev, use_table, da, dt, n, nx, IN, I, sums2,
        partials, ticket);
    TFB_CUDA_CHECK(cudaGetLastError());
    return TFB_OK;
}

extern "C" int tfb_snow_degree_day(tfb_sg P_snow, tfb_sg T_air, tfb_sg c0, tfb_sg T0,
                                   double rho_H2O, double rho_snow, tfb_sg da, double dt,
                                   int32_t ny, int32_t nx, double *SM, double *h_swe,
                                   double *h_snow, double *sums1, double *partials,
                                   uint32_t *ticket, void *stream) {
    TFB_ARG_CHECK(ny > 0 && nx > 0 && SM && h_swe && partials && ticket);
    const int64_t n = (int64_t)ny * nx;
    degree_day_kernel<<<n_blocks(n), kThreads, 0, (cudaStream_t)stream>>>(
        P_snow, T_air, c0, T0, rho_H2O / rho_snow, da, dt, n, nx, SM, h_swe, h_snow, sums1,
        partials, ticket);
    TFB_CUDA_CHECK(cudaGetLastError());
    return TFB_OK;
}

extern "C" int tfb_evap_priestley_taylor(tfb_sg T_air, tfb_sg T_surf, tfb_sg Qn_SW,
                                         tfb_sg Qn_LW, double alpha, double K_soil,
                                         double soil_x, double T_soil_x, tfb_sg da,
                                         double dt, int32_t ny, int32_t nx, double *Qc,
                                         double *ET, double *sums1, double *partials,
                                         uint32_t *ticket, void *stream) {
    TFB_ARG_CHECK(ny > 0 && nx > 0 && ET && partials && ticket);
    const int64_t n = (int64_t)ny * nx;
    priestley_taylor_kernel<<<n_blocks(n), kThreads, 0, (cudaStream_t)stream>>>(
        T_air, T_surf, Qn_SW, Qn_LW, alpha, K_soil, soil_x, T_soil_x, da, dt, n, nx, Qc, ET,
        sums1, partials, ticket);
    TFB_CUDA_CHECK(cudaGetLastError());
    return TFB_OK;
}

extern "C" int64_t tfb_sizeof_met_energy_args(void) { return (int64_t)sizeof(tfb_met_energy_args); }

extern "C" int tfb_met_energy(const tfb_met_energy_args *ap, void *stream) {
    TFB_ARG_CHECK(ap != nullptr && ap->ny > 0 && ap->nx > 0);
    // day terms on the host (solar_funcs.py:147-234), shared by every cell
    const double G = (2.0 * M_PI) * ap->julian_day / 365.0;
    MetDay dy;
    dy.E0 = 1.000110 + (0.034221 * cos(G)) + (0.001280 * sin(G)) + (0.000719 * cos(2.0 * G)) +
            (0.000077 * sin(2.0 * G));
    dy.delta = 0.006918 - (0.399912 * cos(G)) + (0.070257 * sin(G)) - (0.006758 * cos(2.0 * G)) +
               (0.000907 * sin(2.0 * G)) - (0.002697 * cos(3.0 * G)) + (0.001480 * sin(3.0 * G));
    dy.sin_d = sin(dy.delta); dy.cos_d = cos(dy.delta); dy.tan_d = tan(dy.delta);
    const int64_t n = (int64_t)ap->ny * ap->nx;
    met_energy_kernel<<<n_blocks(n), kThreads, 0, (cudaStream_t)stream>>>(*ap, dy);
    TFB_CUDA_CHECK(cudaGetLastError());
    return TFB_OK;
}

extern "C" int tfb_channels_minmax(const double *Q, const double *u, const double *d,
                                   const double *vol, int32_t ny, int32_t nx, double *out6,
                                   double *vol_sum, double *partials, uint32_t *ticket,
                                   void *stream) {
    TFB_ARG_CHECK(Q && u && d && vol && out6 && vol_sum && partials && ticket);
    const int64_t n = (int64_t)ny * nx;
    minmax_kernel<<<n_blocks(n), kThreads, 0, (cudaStream_t)stream>>>(
        Q, u, d, vol, ny, nx, out6, vol_sum, partials, ticket);
    TFB_CUDA_CHECK(cudaGetLastError());
    return TFB_OK;
}

extern "C" int tfb_met_albedo(tfb_sg P_snow, tfb_sg T_air, tfb_sg h_snow, tfb_sg h_ice,
                              int32_t ny, int32_t nx, double dt, double rho_H2O, double rho_snow,
                              int32_t aging, double *ring, int32_t slots, int32_t head,
                              double *n_days, double *albedo, void *stream) {
    TFB_ARG_CHECK(albedo && ny > 0 && nx > 0);
    TFB_ARG_CHECK(!aging || (ring && n_days && slots > 0 && head >= 0 && head < slots));
    const int64_t n = (int64_t)ny * nx;
    const double ratio = rho_H2O / rho_snow;                                      // :2016
    met_albedo_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        P_snow, T_air, h_snow, h_ice, n, ring, slots, head, n_days, dt, ratio, dt / 86400.0,
        aging, albedo);
    TFB_CUDA_CHECK(cudaGetLastError());
    return TFB_OK;
}
