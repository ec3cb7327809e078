// D8 preprocessing toolkit for B200 (sm_100a): flow codes, tie breaking, flat
// linking, parent IDs, flow length/width, total contributing area, slope.
//
// Replaces components/d8_global.py (start_new_d8_codes :181-353,
// break_flow_grid_ties :355-372, link_flats :374-576, update_parent_ID_grid
// :578-650, update_flow_width_grid :873-958, update_flow_length_grid :960-1027,
// update_area_grid :1029-1332, update_slope_grid :1334-1348) and the tie table
// of components/d8_base.py:711-772.  All of it is integer / fp32 work and must
// be BIT-EXACT against the reference; nothing here may be re-associated.
#include <stdlib.h>
#include "tfb_common.cuh"

#include <string.h>

namespace tfb {

// Tie-break table: index = sum of tied codes, value = code RiverTools keeps.
// Data dumped from the reference's get_resolve_array (d8_base.py:711-772);
// pinned by tests (sum 10161, sha256 b0cd03c82342c315...).
static const uint8_t kResolveHost[256] = {
    0x00,0x01,0x02,0x02,0x04,0x01,0x02,0x02,0x08,0x08,0x02,0x02,0x08,0x08,0x02,0x02,
    0x10,0x01,0x02,0x02,0x04,0x04,0x02,0x02,0x08,0x08,0x08,0x02,0x08,0x08,0x08,0x02,
    0x20,0x20,0x02,0x02,0x20,0x20,0x02,0x02,0x08,0x08,0x08,0x02,0x08,0x08,0x08,0x02,
    0x20,0x20,0x20,0x20,0x20,0x20,0x02,0x02,0x08,0x08,0x08,0x08,0x08,0x08,0x08,0x08,
    0x40,0x40,0x02,0x02,0x04,0x01,0x02,0x02,0x08,0x08,0x02,0x02,0x08,0x08,0x02,0x02,
    0x10,0x40,0x02,0x02,0x10,0x01,0x02,0x02,0x08,0x08,0x08,0x02,0x08,0x08,0x08,0x02,
    0x20,0x20,0x20,0x20,0x20,0x20,0x02,0x02,0x20,0x20,0x20,0x02,0x08,0x08,0x08,0x02,
    0x20,0x20,0x20,0x20,0x20,0x20,0x20,0x02,0x20,0x20,0x20,0x20,0x08,0x08,0x08,0x08,
    0x80,0x80,0x80,0x80,0x80,0x80,0x02,0x02,0x08,0x80,0x02,0x02,0x08,0x80,0x02,0x02,
    0x80,0x80,0x80,0x80,0x80,0x80,0x02,0x02,0x08,0x08,0x08,0x02,0x08,0x08,0x08,0x02,
    0x20,0x80,0x80,0x80,0x20,0x80,0x02,0x02,0x20,0x80,0x02,0x02,0x08,0x08,0x08,0x02,
    0x20,0x20,0x20,0x80,0x20,0x20,0x20,0x02,0x20,0x20,0x20,0x08,0x08,0x08,0x08,0x02,
    0x80,0x80,0x80,0x80,0x80,0x80,0x80,0x80,0x80,0x80,0x80,0x80,0x80,0x80,0x02,0x02,
    0x80,0x80,0x80,0x80,0x80,0x80,0x80,0x80,0x08,0x80,0x80,0x80,0x08,0x08,0x08,0x02,
    0x20,0x80,0x80,0x80,0x20,0x80,0x80,0x80,0x20,0x80,0x80,0x80,0x20,0x80,0x02,0x80,
    0x20,0x20,0x20,0x80,0x20,0x20,0x20,0x80,0x20,0x20,0x20,0x20,0x20,0x20,0x08,0x08};

__device__ __constant__ uint8_t kResolve[256];
static bool g_resolve_uploaded[64] = {false};

static int ensure_resolve_uploaded() {
    int dev = 0;
    TFB_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 64 && g_resolve_uploaded[dev]) return TFB_OK;
    TFB_CUDA_CHECK(cudaMemcpyToSymbol(kResolve, kResolveHost, 256));
    if (dev < 64) g_resolve_uploaded[dev] = true;
    return TFB_OK;
}

__device__ __forceinline__ float np_maxf(float a, float b) {
    return (a >= b || a != a) ? a : b;
}

// ---------------------------------------------------------------- start codes
__global__ void __launch_bounds__(256)
d8_codes_kernel(const float *__restrict__ dem, int ny, int nx, int halo, int row0,
                int ny_global, float dx, float dy, float dd, double nodata,
                int link_flats, int16_t *__restrict__ code) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y * blockDim.y + threadIdx.y;
    if (c >= nx || r >= ny) return;
    const int64_t i = (int64_t)r * nx + c;
    const int gr = r + row0;
    int out = 0;
    const bool border = (c == 0 || c == nx - 1 || gr == 0 || gr == ny_global - 1);
    const bool rows_ok = (r - 1 >= -halo) && (r + 1 < ny + halo);
    if (!border && rows_ok) {
        const float z = dem[i];
        const float *up = dem + i - nx, *dn = dem + i + nx;
        // slopes in the DEM's dtype, in the reference's order s1..s8
        const float s1 = (z - up[1]) / dd;    // NE
        const float s2 = (z - dem[i + 1]) / dx;   // E
        const float s3 = (z - dn[1]) / dd;    // SE
        const float s4 = (z - dn[0]) / dy;    // S
        const float s5 = (z - dn[-1]) / dd;   // SW
        const float s6 = (z - dem[i - 1]) / dx;   // W
        const float s7 = (z - up[-1]) / dd;   // NW
        const float s8 = (z - up[0]) / dy;    // N
        float mx = np_maxf(s1, s2);
        mx = np_maxf(mx, s3); mx = np_maxf(mx, s4); mx = np_maxf(mx, s5);
        mx = np_maxf(mx, s6); mx = np_maxf(mx, s7); mx = np_maxf(mx, s8);
        int cd = (s1 == mx) * 1 + (s2 == mx) * 2 + (s3 == mx) * 4 + (s4 == mx) * 8 +
                 (s5 == mx) * 16 + (s6 == mx) * 32 + (s7 == mx) * 64 + (s8 == mx) * 128;
        if (!link_flats) {
            if (mx <= 0.0f) cd = 0;
        } else {
            if (mx == 0.0f) cd = -cd;
            if (mx < 0.0f) cd = -300;
        }
        if ((double)z <= nodata || !isfinite(z)) cd = 0;
        out = cd;
    }
    code[i] = (int16_t)out;
}

__global__ void d8_break_ties_kernel(int16_t *code, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int v = code[i];
    if (v > 0) code[i] = (int16_t)kResolve[v & 255];
}

// ---------------------------------------------------------------- flat linking
__global__ void __launch_bounds__(256)
d8_link_flats_kernel(const int16_t *__restrict__ in, int16_t *__restrict__ out, int ny,
                     int nx, int halo, int row0, int ny_global, int *n_flats,
                     int *n_fixed) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y * blockDim.y + threadIdx.y;
    int is_flat = 0, is_fixed = 0;
    if (c < nx && r < ny) {
        const int64_t i = (int64_t)r * nx + c;
        const int gr = r + row0;
        int v = in[i];
        const bool edge = (c == 0 || c == nx - 1 || gr == 0 || gr == ny_global - 1);
        if (v < 0 && v != -300 && !edge && (r - 1 >= -halo) && (r + 1 < ny + halo)) {
            is_flat = 1;
            const int tied = -v;
            int ready = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int nb = in[i + (int64_t)kNbrDRow[k] * nx + kNbrDCol[k]];
                const int opp = 1 << ((k + 4) & 7);
                // neighbour already drains somewhere, and not back into us
                if (nb > 0 && nb != opp && (tied & (1 << k))) ready |= (1 << k);
            }
            if (ready > 0) {
                v = kResolve[ready];
                is_fixed = 1;
            }
        }
        out[i] = (int16_t)v;
    }
    const unsigned mf = __ballot_sync(0xffffffffu, is_flat);
    const unsigned mx = __ballot_sync(0xffffffffu, is_fixed);
    if ((threadIdx.x & 31) == 0 && (threadIdx.y * blockDim.x + threadIdx.x) % 32 == 0) {
        if (mf) atomicAdd(n_flats, __popc(mf));
        if (mx) atomicAdd(n_fixed, __popc(mx));
    }
}

__global__ void d8_finalize_kernel(const int16_t *__restrict__ in, uint8_t *__restrict__ out,
                                   int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int v = in[i];
    // valid_code_map (d8_base.py:597-607): only the 8 single-direction codes
    const bool ok = (v > 0 && v <= 128 && (v & (v - 1)) == 0);
    out[i] = ok ? (uint8_t)v : (uint8_t)0;
}

// ---------------------------------------------------------------- ds / dw
__global__ void d8_ds_dw_kernel(const uint8_t *__restrict__ code, const float *dx,
                                const float *dy, const float *dd, int ny, int nx,
                                float *ds, float *dw) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y * blockDim.y + threadIdx.y;
    if (c >= nx || r >= ny) return;
    const int64_t i = (int64_t)r * nx + c;
    const unsigned cd = code[i];
    if (ds) ds[i] = ds_of(cd, dx[r], dy[r], dd[r]);
    if (dw) dw[i] = dw_of(cd, dx[r], dy[r], dd[r]);
}

// ---------------------------------------------------------------- parent IDs
__global__ void d8_parent_kernel(const uint8_t *__restrict__ code, int ny, int nx,
                                 int32_t *pid) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y * blockDim.y + threadIdx.y;
    if (c >= nx || r >= ny) return;
    const int64_t i = (int64_t)r * nx + c;
    const unsigned cd = code[i];
    int32_t p = 0;
    if (cd != 0 && c != 0 && c != nx - 1 && r != 0 && r != ny - 1) {
        int dr, dc;
        code_offset(cd, dr, dc);
        p = (int32_t)(i + (int64_t)dr * nx + dc);
    }
    pid[i] = p;
}

// ---------------------------------------------------------------- slope
__global__ void d8_slope_kernel(const float *__restrict__ dem, const uint8_t *__restrict__ code,
                                const float *dx, const float *dy, const float *dd, int ny,
                                int nx, float *S) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y * blockDim.y + threadIdx.y;
    if (c >= nx || r >= ny) return;
    const int64_t i = (int64_t)r * nx + c;
    const unsigned cd = code[i];
    float s = 0.0f;
    if (cd != 0 && c != 0 && c != nx - 1 && r != 0 && r != ny - 1) {
        int dr, dc;
        code_offset(cd, dr, dc);
        const float zp = dem[i + (int64_t)dr * nx + dc];
        s = (dem[i] - zp) / ds_of(cd, dx[r], dy[r], dd[r]);
    }
    S[i] = s;
}

// ---------------------------------------------------------------- area
// number of children = neighbours k whose code points back at the cell
__device__ __forceinline__ int child_mask(const uint8_t *__restrict__ code, int ny, int nx,
                                          int r, int c) {
    int m = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int rr = r + kNbrDRow[k], cc = c + kNbrDCol[k];
        if (rr < 0 || rr >= ny || cc < 0 || cc >= nx) continue;
        if (code[(int64_t)rr * nx + cc] == (uint8_t)(1u << ((k + 4) & 7))) m |= (1 << k);
    }
    return m;
}

// Per-cell word of the area pass (kept in the caller's int32 scratch):
//   bits 0-3  children that have not arrived yet      bits 8-15  child mask (neighbour k
//   bits 16-23 D8 code                                 drains into the cell)
//   bit 24    the cell is the ONLY child of its parent bit 25    the parent is a flow cell
constexpr int kWOnly = 1 << 24, kWParent = 1 << 25;

__global__ void __launch_bounds__(256)
d8_area_init_kernel(const uint8_t *__restrict__ code, int ny, int nx, float *A,
                    int64_t *count, int32_t *word) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y * blockDim.y + threadIdx.y;
    if (c >= nx || r >= ny) return;
    const int64_t i = (int64_t)r * nx + c;
    A[i] = 0.0f;
    if (count) count[i] = 0;
    const unsigned cd = code[i];
    int w = 0;
    if (cd != 0) {
        const int m = child_mask(code, ny, nx, r, c);
        w = __popc(m) | (m << 8) | ((int)cd << 16);
        int dr, dc;
        code_offset(cd, dr, dc);
        const int pr = r + dr, pc = c + dc;
        if (pr >= 0 && pr < ny && pc >= 0 && pc < nx && code[(int64_t)pr * nx + pc] != 0) {
            w |= kWParent;
            if (__popc(child_mask(code, ny, nx, pr, pc)) == 1) w |= kWOnly;
        }
    }
    word[i] = w;
}

// One topological pass: a thread starts at every source cell (flow cell with
// no children) and walks downstream for as long as it is the LAST child to
// arrive at the next cell; each cell's value is pixel_area (stored fp32) plus
// its children's final values in neighbour order k = 0..7, exactly the
// reference's per-cell sum (d8_global.py:1223-1258), so the result does not
// depend on the schedule.  Along unbranched reaches (the cell is its parent's only
// child) no atomic or fence is needed and the running value stays in a register:
// one dependent load per cell instead of fence + atomic + reload.
__global__ void __launch_bounds__(256)
d8_area_walk_kernel(const uint8_t *__restrict__ code, int ny, int nx, double pixel_area,
                    const double *__restrict__ pa_row, float *A, int64_t *count,
                    int32_t *word, unsigned long long *n_done) {
    const int c0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int r0 = blockIdx.y * blockDim.y + threadIdx.y;
    if (c0 >= nx || r0 >= ny) return;
    int r = r0, c = c0;
    int64_t i = (int64_t)r * nx + c;
    int w = word[i];
    if (((w >> 16) & 0xff) == 0 || ((w >> 8) & 0xff) != 0) return;    // not a source
    unsigned long long done = 0;
    volatile float *Av = A;
    volatile int64_t *Cv = count;
    float a = (float)(pa_row ? pa_row[r] : pixel_area);
    int64_t n = 1;
    while (true) {
        Av[i] = a;
        if (count) Cv[i] = n;
        ++done;
        if (!(w & kWParent)) break;       // drains off the grid or into a noflow cell
        int dr, dc;
        code_offset((unsigned)((w >> 16) & 0xff), dr, dc);
        r += dr; c += dc;
        const int64_t ip = (int64_t)r * nx + c;
        const float pa = (float)(pa_row ? pa_row[r] : pixel_area);
        if (w & kWOnly) {
            w = __ldcg(word + ip);        // static bits only: nobody else touches this word
            a = pa + a;                   // the single child's value is still in the register
            n = 1 + n;
        } else {
            __threadfence();              // publish A[i] before announcing arrival
            const int old = atomicSub(&word[ip], 1);
            if ((old & 0xf) != 1) break;  // a sibling arrives later and continues
            __threadfence();
            w = old;
            const int cm = (old >> 8) & 0xff;
            a = pa;
            n = 1;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (cm & (1 << k)) {
                    const int64_t j = ip + (int64_t)kNbrDRow[k] * nx + kNbrDCol[k];
                    a += Av[j];
                    if (count) n += Cv[j];
                }
            }
        }
        i = ip;
    }
    atomicAdd(n_done, done);
}

static dim3 grid2d(int ny, int nx, dim3 b) {
    return dim3((nx + b.x - 1) / b.x, (ny + b.y - 1) / b.y);
}

}  // namespace tfb

using namespace tfb;

extern "C" int tfb_d8_resolve_table(uint8_t *out256) {
    TFB_ARG_CHECK(out256 != nullptr);
    memcpy(out256, kResolveHost, 256);
    return TFB_OK;
}

extern "C" int tfb_d8_codes(const float *dem, int32_t ny, int32_t nx, int32_t halo,
                            int32_t row0, int32_t ny_global, float dx, float dy, float dd,
                            double nodata, int32_t link_flats, int16_t *code,
                            void *stream) {
    TFB_ARG_CHECK(dem && code && ny > 0 && nx > 0 && halo >= 0);
    TFB_ARG_CHECK(row0 >= 0 && row0 + ny <= ny_global);
    dim3 b(32, 8);
    d8_codes_kernel<<<grid2d(ny, nx, b), b, 0, (cudaStream_t)stream>>>(
        dem, ny, nx, halo, row0, ny_global, dx, dy, dd, nodata, link_flats, code);
    TFB_CUDA_CHECK(cudaGetLastError());
    return TFB_OK;
}

extern "C" int tfb_d8_break_ties(int16_t *code, int64_t n, void *stream) {
    TFB_ARG_CHECK(code && n > 0);
    int rc = ensure_resolve_uploaded();
    if (rc) return rc;
    d8_break_ties_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(code, n);
    TFB_CUDA_CHECK(cudaGetLastError());
    return TFB_OK;
}

extern "C" int tfb_d8_link_flats_sweep(const int16_t *code_in, int16_t *code_out,
                                       int32_t ny, int32_t nx, int32_t halo, int32_t row0,
                                       int32_t ny_global, int32_t *n_flats,
                                       int32_t *n_fixed, void *stream) {
    TFB_ARG_CHECK(code_in && code_out && code_in != code_out && n_flats && n_fixed);
    TFB_ARG_CHECK(ny > 0 && nx > 0 && halo >= 0);
    int rc = ensure_resolve_uploaded();
    if (rc) return rc;
    dim3 b(32, 8);
    d8_link_flats_kernel<<<grid2d(ny, nx, b), b, 0, (cudaStream_t)stream>>>(
        code_in, code_out, ny, nx, halo, row0, ny_global, n_flats, n_fixed);
    TFB_CUDA_CHECK(cudaGetLastError());
    return TFB_OK;
}

extern "C" int tfb_d8_link_flats(int16_t *code, int16_t *scratch, int32_t ny, int32_t nx,
                                 int32_t max_reps, int32_t *n_reps, void *stream) {
    TFB_ARG_CHECK(code && scratch && ny > 0 && nx > 0);
    cudaStream_t st = (cudaStream_t)stream;
    int32_t *d_cnt = (int32_t *)small_scratch();
    TFB_ARG_CHECK(d_cnt != nullptr);
    int16_t *cur = code, *nxt = scratch;
    int reps = 0, rc = TFB_OK;
    while (true) {
        ++reps;
        int32_t h[2] = {0, 0};
        cudaError_t e = cudaMemsetAsync(d_cnt, 0, 2 * sizeof(int32_t), st);
        if (e == cudaSuccess) {
            rc = tfb_d8_link_flats_sweep(cur, nxt, ny, nx, 0, 0, ny, d_cnt, d_cnt + 1, stream);
            if (rc) break;
            e = cudaMemcpyAsync(h, d_cnt, sizeof(h), cudaMemcpyDeviceToHost, st);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            set_error("tfb_d8_link_flats: %s", cudaGetErrorString(e));
            rc = TFB_ERR_CUDA;
            break;
        }
        if (h[0] == 0) break;               // no flats left (sweep was a copy)
        int16_t *t = cur; cur = nxt; nxt = t;   // Jacobi: results live in `nxt`
        if (h[1] == h[0] || h[1] == 0) break;   // all fixed, or none fixable
        if (max_reps > 0 && reps >= max_reps) { rc = TFB_ERR_NOT_CONVERGED; break; }
    }
    if (rc == TFB_OK && cur != code) {
        cudaError_t e = cudaMemcpyAsync(code, cur, (size_t)ny * nx * sizeof(int16_t),
                                        cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { set_error("tfb_d8_link_flats: %s", cudaGetErrorString(e)); rc = TFB_ERR_CUDA; }
    }
    if (n_reps) *n_reps = reps;
    return rc;
}

extern "C" int tfb_d8_finalize_codes(const int16_t *code, uint8_t *code_u8, int64_t n,
                                     void *stream) {
    TFB_ARG_CHECK(code && code_u8 && n > 0);
    d8_finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(code, code_u8, n);
    TFB_CUDA_CHECK(cudaGetLastError());
    return TFB_OK;
}

extern "C" int tfb_d8_ds_dw(const uint8_t *code, const float *dx, const float *dy,
                            const float *dd, int32_t ny, int32_t nx, float *ds, float *dw,
                            void *stream) {
    TFB_ARG_CHECK(code && dx && dy && dd && ny > 0 && nx > 0 && (ds || dw));
    dim3 b(32, 8);
    d8_ds_dw_kernel<<<grid2d(ny, nx, b), b, 0, (cudaStream_t)stream>>>(code, dx, dy, dd, ny, nx, ds, dw);
    TFB_CUDA_CHECK(cudaGetLastError());
    return TFB_OK;
}

extern "C" int tfb_d8_parent_ids(const uint8_t *code, int32_t ny, int32_t nx, int32_t *pid,
                                 void *stream) {
    TFB_ARG_CHECK(code && pid && ny > 0 && nx > 0);
    TFB_ARG_CHECK((int64_t)ny * nx < 2147483647LL);
    dim3 b(32, 8);
    d8_parent_kernel<<<grid2d(ny, nx, b), b, 0, (cudaStream_t)stream>>>(code, ny, nx, pid);
    TFB_CUDA_CHECK(cudaGetLastError());
    return TFB_OK;
}

extern "C" int tfb_d8_slope(const float *dem, const uint8_t *code, const float *dx,
                            const float *dy, const float *dd, int32_t ny, int32_t nx,
                            float *S, void *stream) {
    TFB_ARG_CHECK(dem && code && dx && dy && dd && S && ny > 0 && nx > 0);
    dim3 b(32, 8);
    d8_slope_kernel<<<grid2d(ny, nx, b), b, 0, (cudaStream_t)stream>>>(dem, code, dx, dy, dd, ny, nx, S);
    TFB_CUDA_CHECK(cudaGetLastError());
    return TFB_OK;
}

// ---------------------------------------------------------------- aspect
struct AspectLut { float a[9]; };
__global__ void d8_aspect_kernel(const uint8_t *__restrict__ code, int64_t n, AspectLut lut,
                                 float *__restrict__ aspect) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned cd = code[i];
    // Jenson codes -> index into linspace(0, 2 pi, 9) (d8_global.py:1358-1373)
    int k = 0;
    switch (cd) {
        case 1: k = 1; break;   case 2: k = 0; break;   case 4: k = 7; break;
        case 8: k = 6; break;   case 16: k = 5; break;  case 32: k = 4; break;
        case 64: k = 3; break;  case 128: k = 2; break; default: k = 0;
    }
    aspect[i] = (cd == 0) ? 0.0f : lut.a[k];
}

extern "C" int tfb_d8_aspect(const uint8_t *code, int64_t n, float *aspect, void *stream) {
    TFB_ARG_CHECK(code && aspect && n > 0);
    AspectLut lut;
    const double step = (2.0 * 3.141592653589793 - 0.0) / 8.0;       // np.linspace
    for (int k = 0; k < 9; ++k) lut.a[k] = (float)(k == 8 ? 2.0 * 3.141592653589793 : k * step);
    d8_aspect_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(code, n, lut, aspect);
    TFB_CUDA_CHECK(cudaGetLastError());
    return TFB_OK;
}

// ---------------------------------------------------------------- new slopes
// utils/new_slopes.get_new_slope_grid (:174-316): every cell walks down its D8 flow
// path until the drop from ITS OWN elevation is positive; slope = drop / path length.
// The reference advances all cells together for J iterations, J = the longest flow
// path of the grid; walkers that have passed their noflow cell sit on cell (0, 0)
// (parent ID 0) and are compared with DEM[0, 0] there -- reproduced below.
//
// Longest flow path: the same last-child-continues walk as the area pass, carrying
// the number of cells on the longest upstream chain.
__global__ void __launch_bounds__(256)
d8_longest_walk_kernel(int ny, int nx, int32_t *len, const int32_t *__restrict__ word, int32_t *J) {
    const int c0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int r0 = blockIdx.y * blockDim.y + threadIdx.y;
    if (c0 >= nx || r0 >= ny) return;
    int r = r0, c = c0;
    int64_t i = (int64_t)r * nx + c;
    int w = __ldg(word + i);
    if (((w >> 16) & 0xff) == 0 || ((w >> 8) & 0xff) != 0) return;    // not a source
    int h = 1;                        // cells on the longest chain that ends here (0 = unknown)
    while (true) {
        __stcg(len + i, h);
        if (!(w & kWParent)) {                             // the next cell is a noflow cell
            // parent ID 0 is the reference's "no parent" marker AND the ID of cell (0, 0):
            // a path through (1, 1) -> NW already counts as finished on (1, 1)
            if (r == 1 && c == 1 && ((w >> 16) & 0xff) == 64 && h > 1) h -= 1;
            atomicMax(J, h);
            break;
        }
        int dr, dc;
        code_offset((unsigned)((w >> 16) & 0xff), dr, dc);
        r += dr; c += dc;
        const int64_t ip = (int64_t)r * nx + c;
        const int wp = __ldg(word + ip);
        if (w & kWOnly) {
            h += 1;
        } else {
            // same protocol as d8_area_walk2_kernel: store mine, fence.sc, read all siblings
            asm volatile("fence.sc.gpu;" ::: "memory");
            const int cm = (wp >> 8) & 0xff;
            int best = 0;
            bool ready = true;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (cm & (1 << k)) {
                    const int v = __ldcg(len + ip + (int64_t)kNbrDRow[k] * nx + kNbrDCol[k]);
                    ready = ready && (v != 0);
                    best = max(best, v);
                }
            }
            if (!ready) break;
            h = 1 + best;
        }
        w = wp;
        i = ip;
    }
}

__global__ void __launch_bounds__(256)
d8_new_slope_kernel(const float *__restrict__ dem, const uint8_t *__restrict__ code,
                    const float *__restrict__ dx, const float *__restrict__ dy,
                    const float *__restrict__ dd, int ny, int nx, float nodata,
                    const int32_t *__restrict__ Jp, double *__restrict__ slope) {
    const int c0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int r0 = blockIdx.y * blockDim.y + threadIdx.y;
    if (c0 >= nx || r0 >= ny) return;
    const int64_t i0 = (int64_t)r0 * nx + c0;
    unsigned cd = code[i0];
    double s = 0.0;
    if (cd != 0) {                                   // own code 0: dz is forced to 0 (:221)
        // iterations the reference executes: J, capped by n_reps > 4 nx (:255)
        const int64_t cap = 4LL * nx + 1;
        const int64_t J = (*Jp < 1) ? 1 : ((*Jp < cap) ? *Jp : cap);
        const float z2 = dem[i0];
        int r = r0, c = c0;
        double dist = (double)ds_of(cd, dx[r], dy[r], dd[r]);
        for (int64_t j = 1; j <= J; ++j) {
            int dr = 0, dc = 0;
            if (cd != 0) { code_offset(cd, dr, dc); r += dr; c += dc; }
            else { r = 0; c = 0; }                   // parent ID of a noflow cell is 0
            const float z1 = dem[(int64_t)r * nx + c];
            if (!(z1 <= nodata)) {
                const float dz = __fsub_rn(z2, z1);
                if (dz > 0.0f) { s = (double)dz / dist; break; }
            }
            const bool was_noflow = (cd == 0);
            cd = code[(int64_t)r * nx + c];
            if (was_noflow) break;                   // on (0, 0): nothing changes any more
            dist += (double)ds_of(cd, dx[r], dy[r], dd[r]);
        }
    }
    slope[i0] = s;
}

extern "C" int tfb_d8_new_slope(const float *dem, const uint8_t *code, const float *dx,
                                const float *dy, const float *dd, int32_t ny, int32_t nx,
                                float nodata, int32_t *scratch, double *slope,
                                int32_t *longest_path, void *stream) {
    TFB_ARG_CHECK(dem && code && dx && dy && dd && scratch && slope && ny > 0 && nx > 0);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = (int64_t)ny * nx;
    int32_t *word = scratch, *len = scratch + n, *J = scratch + 2 * n;
    dim3 b(32, 8);
    // the area pass's init kernel zeroes its fp32 output: used here for `len` (+ J)
    TFB_CUDA_CHECK(cudaMemsetAsync(J, 0, sizeof(int32_t), st));
    d8_area_init_kernel<<<grid2d(ny, nx, b), b, 0, st>>>(code, ny, nx, (float *)len, nullptr, word);
    d8_longest_walk_kernel<<<grid2d(ny, nx, b), b, 0, st>>>(ny, nx, len, word, J);
    d8_new_slope_kernel<<<grid2d(ny, nx, b), b, 0, st>>>(dem, code, dx, dy, dd, ny, nx, nodata,
                                                        J, slope);
    TFB_CUDA_CHECK(cudaGetLastError());
    if (longest_path) {
        TFB_CUDA_CHECK(cudaMemcpyAsync(longest_path, J, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        TFB_CUDA_CHECK(cudaStreamSynchronize(st));
        if (*longest_path < 1) *longest_path = 1;
    }
    return TFB_OK;
}

// Second form of the topological pass, without atomics.  Every cell owns one 64-bit word
// {fp32 area, uint32 count}, 0 = not computed yet (a computed count is >= 1).  A walker
// stores its cell's word, issues fence.sc and then reads the words of ALL children of the
// next cell: if one is still 0 it stops (that child's walker will come by later), else
// it sums them in the reference's order and goes on.  With sequentially consistent
// fences between "store mine" and "load the siblings'" at least one of any two siblings
// sees the other (the store-buffering litmus), so the last arrival always continues;
// when two arrive within a fence latency of each other BOTH may continue -- they compute
// and store identical words, which is harmless.  Per junction the critical path is one
// fence and one round trip to L2 instead of fence + atomic + fence + loads.
__device__ __forceinline__ unsigned long long pack_an(float a, unsigned n) {
    return ((unsigned long long)n << 32) | (unsigned long long)__float_as_uint(a);
}

__global__ void __launch_bounds__(256)
d8_area_walk2_kernel(int ny, int nx, double pixel_area, const double *__restrict__ pa_row,
                     unsigned long long *cell, const int32_t *__restrict__ word) {
    const int c0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int r0 = blockIdx.y * blockDim.y + threadIdx.y;
    if (c0 >= nx || r0 >= ny) return;
    int r = r0, c = c0;
    int64_t i = (int64_t)r * nx + c;
    int w = __ldg(word + i);
    if (((w >> 16) & 0xff) == 0 || ((w >> 8) & 0xff) != 0) return;    // not a source
    float a = (float)(pa_row ? pa_row[r] : pixel_area);
    unsigned n = 1u;
    while (true) {
        __stcg(cell + i, pack_an(a, n));
        if (!(w & kWParent)) break;       // drains into a noflow cell
        int dr, dc;
        code_offset((unsigned)((w >> 16) & 0xff), dr, dc);
        r += dr; c += dc;
        const int64_t ip = (int64_t)r * nx + c;
        const float pa = (float)(pa_row ? pa_row[r] : pixel_area);
        const int wp = __ldg(word + ip);
        if (w & kWOnly) {
            a = pa + a;                   // the single child's value is still in the register
            n = 1u + n;
        } else {
            asm volatile("fence.sc.gpu;" ::: "memory");
            const int cm = (wp >> 8) & 0xff;
            unsigned long long v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k)
                v[k] = (cm & (1 << k))
                           ? __ldcg(cell + ip + (int64_t)kNbrDRow[k] * nx + kNbrDCol[k]) : 1ull;
            bool ready = true;
#pragma unroll
            for (int k = 0; k < 8; ++k) ready = ready && (v[k] != 0ull);
            if (!ready) break;            // a sibling arrives later and continues
            a = pa;
            n = 1u;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (cm & (1 << k)) {
                    a += __uint_as_float((unsigned)(v[k] & 0xffffffffull));
                    n += (unsigned)(v[k] >> 32);
                }
            }
        }
        w = wp;
        i = ip;
    }
}

// ---- contributing area across row slabs (one process per GPU) ---------------------------
// Same words, same last-child-continues walk, but a slab only owns rows [0, ny) of the
// global grid and sees one halo row of D8 codes and of cell words on either side.  A walker
// stops when its next cell belongs to the neighbour; the boundary rows of the words are then
// exchanged (host: NCCL send/recv of two rows) and `resume` lets every halo cell whose word
// has arrived act as the walker that has just stored it.  Rounds repeat until no boundary
// row changes on any rank: one round per crossing of a slab boundary along the longest
// dependency chain.  Per-cell sums keep the reference's neighbour order, so areas and
// counts are bit-identical to the single-GPU pass (d8_global.update_area_grid :1029-1332).
__global__ void __launch_bounds__(256)
d8_area_slab_init_kernel(const uint8_t *__restrict__ code, int ny, int nx, int row0, int ny_global,
                         unsigned long long *cell, int32_t *word) {
    // code / cell / word point at the first OWNED row; rows -1 and ny are the halo
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = (int)(blockIdx.y * blockDim.y + threadIdx.y) - 1;
    if (c >= nx || r > ny) return;
    const int64_t i = (int64_t)r * nx + c;
    cell[i] = 0ull;
    const int lo = (row0 == 0) ? 0 : -1, hi = (row0 + ny == ny_global) ? ny : ny + 1;   // readable rows
    auto cmask = [&](int rr0, int cc0) {
        int m = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int rr = rr0 + kNbrDRow[k], cc = cc0 + kNbrDCol[k];
            if (rr < lo || rr >= hi || cc < 0 || cc >= nx) continue;
            if (code[(int64_t)rr * nx + cc] == (uint8_t)(1u << ((k + 4) & 7))) m |= (1 << k);
        }
        return m;
    };
    const bool in_grid = (r >= lo && r < hi);
    const unsigned cd = in_grid ? code[i] : 0u;
    int w = 0;
    if (cd != 0) {
        int dr, dc;
        code_offset(cd, dr, dc);
        const int pr = r + dr, pc = c + dc;
        const bool owned = (r >= 0 && r < ny);
        const int m = owned ? cmask(r, c) : 0;           // a halo cell's own children are not ours
        w = __popc(m) | (m << 8) | ((int)cd << 16);
        if (pr >= lo && pr < hi && pc >= 0 && pc < nx && code[(int64_t)pr * nx + pc] != 0) {
            w |= kWParent;
            // "only child" needs the parent's full neighbourhood: parents in owned rows
            if (pr >= 0 && pr < ny && owned && __popc(cmask(pr, pc)) == 1) w |= kWOnly;
        }
    }
    word[i] = w;
}

// the walk of d8_area_walk2_kernel from cell (r, c) holding {a, n}; `stored`: its word is
// already in memory (a halo cell that arrived with the exchange)
__device__ __forceinline__ void area_walk_slab(int ny, int nx, double pixel_area,
                                               const double *__restrict__ pa_row,
                                               unsigned long long *cell,
                                               const int32_t *__restrict__ word, int r, int c,
                                               int w, float a, unsigned n, bool stored) {
    int64_t i = (int64_t)r * nx + c;
    while (true) {
        if (!stored) __stcg(cell + i, pack_an(a, n));
        if (!(w & kWParent)) break;       // drains into a noflow cell
        int dr, dc;
        code_offset((unsigned)((w >> 16) & 0xff), dr, dc);
        r += dr; c += dc;
        if (r < 0 || r >= ny) break;      // the next cell is the neighbour's: it resumes there
        const int64_t ip = (int64_t)r * nx + c;
        const float pa = (float)(pa_row ? pa_row[r] : pixel_area);
        const int wp = __ldg(word + ip);
        if ((w & kWOnly) && !stored) {
            a = pa + a;
            n = 1u + n;
        } else {
            asm volatile("fence.sc.gpu;" ::: "memory");
            const int cm = (wp >> 8) & 0xff;
            unsigned long long v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k)
                v[k] = (cm & (1 << k))
                           ? __ldcg(cell + ip + (int64_t)kNbrDRow[k] * nx + kNbrDCol[k]) : 1ull;
            bool ready = true;
#pragma unroll
            for (int k = 0; k < 8; ++k) ready = ready && (v[k] != 0ull);
            if (!ready) break;            // a sibling arrives later (perhaps in a later round)
            if (stored && __ldcg(cell + ip) != 0ull) break;    // resumed before: already done
            a = pa;
            n = 1u;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (cm & (1 << k)) {
                    a += __uint_as_float((unsigned)(v[k] & 0xffffffffull));
                    n += (unsigned)(v[k] >> 32);
                }
            }
        }
        stored = false;
        w = wp;
        i = ip;
    }
}

__global__ void __launch_bounds__(256)
d8_area_slab_walk_kernel(int ny, int nx, double pixel_area, const double *__restrict__ pa_row,
                         unsigned long long *cell, const int32_t *__restrict__ word) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y * blockDim.y + threadIdx.y;
    if (c >= nx || r >= ny) return;
    const int w = __ldg(word + (int64_t)r * nx + c);
    if (((w >> 16) & 0xff) == 0 || ((w >> 8) & 0xff) != 0) return;    // not a source
    area_walk_slab(ny, nx, pixel_area, pa_row, cell, word, r, c, w,
                   (float)(pa_row ? pa_row[r] : pixel_area), 1u, false);
}

__global__ void __launch_bounds__(256)
d8_area_slab_resume_kernel(int ny, int nx, double pixel_area, const double *__restrict__ pa_row,
                           unsigned long long *cell, const int32_t *__restrict__ word) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nx) return;
    const int r = (blockIdx.y == 0) ? -1 : ny;                        // the two halo rows
    const int64_t i = (int64_t)r * nx + c;
    const int w = __ldg(word + i);
    if (!(w & kWParent)) return;
    const unsigned long long v = __ldcg(cell + i);
    if (v == 0ull) return;                                            // not arrived yet
    area_walk_slab(ny, nx, pixel_area, pa_row, cell, word, r, c, w,
                   __uint_as_float((unsigned)(v & 0xffffffffull)), (unsigned)(v >> 32), true);
}

__global__ void __launch_bounds__(256)
d8_area_unpack_kernel(const unsigned long long *cell, int64_t ncell, float *__restrict__ A,
                      int64_t *count, unsigned long long *n_done) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool done = false;
    if (i < ncell) {
        const unsigned long long v = cell[i];          // may alias count[i]: read, then write
        done = (v != 0ull);
        A[i] = __uint_as_float((unsigned)(v & 0xffffffffull));
        if (count) count[i] = (int64_t)(v >> 32);
    }
    const unsigned m = __ballot_sync(0xffffffffu, done);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_done, (unsigned long long)__popc(m));
}

extern "C" int tfb_d8_area(const uint8_t *code, int32_t ny, int32_t nx, double pixel_area,
                           const double *pixel_area_row, float *A, int64_t *count,
                           int32_t *pending, int64_t *n_done, void *stream) {
    TFB_ARG_CHECK(code && A && pending && ny > 0 && nx > 0);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long *d_done = (unsigned long long *)small_scratch();
    TFB_ARG_CHECK(d_done != nullptr);
    cudaError_t e = cudaMemsetAsync(d_done, 0, sizeof(unsigned long long), st);
    dim3 b(32, 8);
    unsigned long long *tmp = nullptr;
    const bool atomic_form = getenv("TFB_D8_AREA_ATOMIC") != nullptr ||
                             (int64_t)ny * nx >= (1LL << 32);       // counts must fit 32 bits
    if (e == cudaSuccess && atomic_form) {
        d8_area_init_kernel<<<grid2d(ny, nx, b), b, 0, st>>>(code, ny, nx, A, count, pending);
        d8_area_walk_kernel<<<grid2d(ny, nx, b), b, 0, st>>>(code, ny, nx, pixel_area,
                                                            pixel_area_row, A, count,
                                                            pending, d_done);
        e = cudaGetLastError();
    } else if (e == cudaSuccess) {
        const int64_t n = (int64_t)ny * nx;
        unsigned long long *cell = (unsigned long long *)count;     // the words live in `count`
        if (!cell) {
            e = cudaMalloc(&tmp, (size_t)n * sizeof(unsigned long long));
            cell = tmp;
        }
        if (e == cudaSuccess) {
            // init zeroes A (and count = the words); without `count` the words are cleared here
            d8_area_init_kernel<<<grid2d(ny, nx, b), b, 0, st>>>(code, ny, nx, A, count, pending);
            if (tmp) e = cudaMemsetAsync(tmp, 0, (size_t)n * sizeof(unsigned long long), st);
        }
        if (e == cudaSuccess) {
            d8_area_walk2_kernel<<<grid2d(ny, nx, b), b, 0, st>>>(ny, nx, pixel_area, pixel_area_row,
                                                                 cell, pending);
            d8_area_unpack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(cell, n, A, count, d_done);
            e = cudaGetLastError();
        }
    }
    unsigned long long h = 0;
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(&h, d_done, sizeof(h), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (tmp) cudaFree(tmp);
    if (e != cudaSuccess) {
        set_error("tfb_d8_area: %s", cudaGetErrorString(e));
        return TFB_ERR_CUDA;
    }
    if (n_done) *n_done = (int64_t)h;
    return TFB_OK;
}

// ---- row-slab form of the area pass (see d8_area_slab_init_kernel) ---------------------------
extern "C" int tfb_d8_area_slab_init(const uint8_t *code, int32_t ny, int32_t nx, int32_t row0,
                                     int32_t ny_global, uint64_t *cell, int32_t *word,
                                     void *stream) {
    TFB_ARG_CHECK(code && cell && word && ny > 0 && nx > 0 && row0 >= 0 && row0 + ny <= ny_global);
    dim3 b(32, 8), g((nx + 31) / 32, (ny + 2 + 7) / 8);
    d8_area_slab_init_kernel<<<g, b, 0, (cudaStream_t)stream>>>(code, ny, nx, row0, ny_global,
                                                               (unsigned long long *)cell, word);
    TFB_CUDA_CHECK(cudaGetLastError());
    return TFB_OK;
}

extern "C" int tfb_d8_area_slab_walk(int32_t ny, int32_t nx, double pixel_area,
                                     const double *pixel_area_row, uint64_t *cell,
                                     const int32_t *word, int32_t resume, void *stream) {
    TFB_ARG_CHECK(cell && word && ny > 0 && nx > 0);
    cudaStream_t st = (cudaStream_t)stream;
    if (!resume) {
        dim3 b(32, 8);
        d8_area_slab_walk_kernel<<<grid2d(ny, nx, b), b, 0, st>>>(
            ny, nx, pixel_area, pixel_area_row, (unsigned long long *)cell, word);
    } else {
        dim3 b(256, 1), g((nx + 255) / 256, 2);
        d8_area_slab_resume_kernel<<<g, b, 0, st>>>(ny, nx, pixel_area, pixel_area_row,
                                                    (unsigned long long *)cell, word);
    }
    TFB_CUDA_CHECK(cudaGetLastError());
    return TFB_OK;
}

extern "C" int tfb_d8_area_unpack(const uint64_t *cell, int64_t n, float *A, int64_t *count,
                                  int64_t *n_done, void *stream) {
    TFB_ARG_CHECK(cell && A && n > 0);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long *d_done = (unsigned long long *)small_scratch();
    TFB_ARG_CHECK(d_done != nullptr);
    TFB_CUDA_CHECK(cudaMemsetAsync(d_done, 0, sizeof(unsigned long long), st));
    d8_area_unpack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
        (const unsigned long long *)cell, n, A, count, d_done);
    TFB_CUDA_CHECK(cudaGetLastError());
    unsigned long long h = 0;
    TFB_CUDA_CHECK(cudaMemcpyAsync(&h, d_done, sizeof(h), cudaMemcpyDeviceToHost, st));
    TFB_CUDA_CHECK(cudaStreamSynchronize(st));
    if (n_done) *n_done = (int64_t)h;
    return TFB_OK;
}
