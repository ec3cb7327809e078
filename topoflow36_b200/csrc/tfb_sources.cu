// Stand-alone source-term kernels: one per reference component update().
// Used when the source components run as separate BMI components (EMELI-style
// coupling, each at its own dt); the routing kernel can alternatively evaluate
// the same expressions fused (TFB_CH_SRC_* flags in tfb_channels.cu).
//
// Each kernel is a single coalesced pass over the grid with a deterministic
// block-partial -> last-block reduction for the volume integrals.
#include "tfb_common.cuh"

namespace tfb {

constexpr int kThreads = 256;
constexpr int kMaxSums = 4;

// grid-stride element loop + deterministic reduction of up to kMaxSums sums
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

static int n_blocks(int64_t n) {
    int64_t want = (n + kThreads - 1) / kThreads;
    int64_t cap = (int64_t)sm_count() * 8;      // 8 resident CTAs of 256 per SM
    return (int)(want < cap ? want : cap);
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
