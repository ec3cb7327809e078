// DEM / channel-geometry preparation around the D8 toolkit:
//
//   * depression filling  -- utils/fill_pits.py fill_pits() :358-640 (the priority-flood
//     of Wang & Liu 2006 with a heap) and its seed rule get_start_pixels() :178-356;
//   * power-law grids from contributing area -- utils/parameterize.py
//     get_grid_from_TCA() :93-215 (channel width, Manning n, bankfull depth).
//
// Depression filling on the GPU.  The priority-flood result is the unique solution of
//
//        W(c) = z(c)                                       for seed cells
//        W(c) = max( z(c), min over the 8 neighbours W )   for open cells
//        closed cells are never entered (W = +inf, impassable)
//
// Which cells are closed: fill_pits.py:404-409 takes where(DEM <= nodata) of the 2-D grid,
// keeps only the ROW numbers and uses them as FLAT cell IDs.  So in the reference the
// nodata cells themselves stay open (they lie lowest and are flooded up to their rim like
// any depression) and the cells with flat index r < ny, r a row that holds nodata, are
// walled off instead unless they are seeds.  close_nodata = 0 reproduces that, bit for
// bit; close_nodata = 1 is the documented intent (nodata cells closed).
//
// (the lowest surface >= z on which every cell has a non-ascending path to a seed).  Only
// max / min of input values occur, so ANY order of relaxations started from W = +inf
// converges to bit-identical values.  One warp owns a 32x32 tile in shared memory and
// runs directional Gauss-Seidel sweeps (a sweep carries information across the whole
// tile) until the tile is stable; tiles whose rim changed mark their neighbours dirty
// for the next round; rounds repeat until no tile changed.
#include <math.h>
#include <string.h>
#include "tfb_common.cuh"

namespace tfb {

constexpr int FT = 32;           // tile edge
constexpr int FSW = FT + 3;      // row stride of the W tile (34 used; odd: no bank conflicts)
constexpr int FSZ = FT + 1;      // row stride of the z tile
constexpr int FWARPS = 8;        // tiles per CTA
constexpr int FSMEM = FWARPS * ((FT + 2) * FSW + FT * FSZ) * (int)sizeof(float);

__device__ __forceinline__ float fix_bad(float v, float nodata) {
    return (isfinite(v)) ? v : nodata;                         // fill_pits.py:232-235
}

// row_flag[r] = 1 when row r holds a value <= nodata (after the non-finite fix)
__global__ void __launch_bounds__(256)
fill_rowflag_kernel(const float *__restrict__ dem, int nx, float nodata, uint8_t *row_flag) {
    const int r = blockIdx.x;
    int hit = 0;
    for (int c = threadIdx.x; c < nx; c += blockDim.x)
        hit |= (fix_bad(dem[(int64_t)r * nx + c], nodata) <= nodata);
    hit = __syncthreads_or(hit);
    if (threadIdx.x == 0) row_flag[r] = (uint8_t)(hit != 0);
}

// seeds, closed cells, initial surface
__global__ void __launch_bounds__(256)
fill_init_kernel(float *dem, float *__restrict__ zz, float *__restrict__ W, int ny, int nx,
                 float nodata, float closed_code, int close_nodata,
                 const uint8_t *__restrict__ row_flag) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y * blockDim.y + threadIdx.y;
    if (c >= nx || r >= ny) return;
    const int64_t i = (int64_t)r * nx + c;
    const float raw = dem[i];
    const float z = fix_bad(raw, nodata);
    bool seed;
    if (r == 0 || r == ny - 1 || c == 0 || c == nx - 1) {
        seed = z > nodata;                                     // :240-268
    } else {
        // :278-302 -- the 8 neighbours are np.roll()s of the INTERIOR sub-array, so on the
        // interior's rim they wrap to the opposite interior rim, not to the border cells
        const int my = ny - 2, mx = nx - 2;
        float zs = INFINITY;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            int rr = r - 1 + kNbrDRow[k], cc = c - 1 + kNbrDCol[k];
            rr = (rr < 0) ? rr + my : ((rr >= my) ? rr - my : rr);
            cc = (cc < 0) ? cc + mx : ((cc >= mx) ? cc - mx : cc);
            zs = fminf(zs, fix_bad(dem[(int64_t)(rr + 1) * nx + cc + 1], nodata));
        }
        seed = (z == closed_code) || (z > nodata && zs <= nodata);
    }
    const bool closed = !seed && (close_nodata ? (z <= nodata)               // :404-409, :437
                                               : (i < ny && row_flag[i] != 0));
    zz[i] = closed ? INFINITY : z;
    W[i] = seed ? z : INFINITY;
    if (z != raw) dem[i] = z;                                  // NaN / Inf -> nodata, in place
}

// One round: the warps of a persistent grid pop tiles from this round's work list; a tile
// whose rim changed appends the neighbours that read that rim to the next round's list
// (`queued[t]` = tile t sits in a list and has not been popped yet).
__global__ void __launch_bounds__(32 * FWARPS)
fill_relax_kernel(const float *__restrict__ zz, float *W, int ny, int nx, int tiles_y,
                  int tiles_x, const int *__restrict__ list_cur, int *list_next,
                  unsigned *queued, const unsigned *n_cur, unsigned *n_next, unsigned *pop) {
    extern __shared__ float fsm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *sW = fsm + warp * ((FT + 2) * FSW + FT * FSZ);
    float *sZ = sW + (FT + 2) * FSW;
    const unsigned n_tiles_round = *n_cur;
    while (true) {
    unsigned idx = 0;
    if (lane == 0) idx = atomicAdd(pop, 1u);
    idx = __shfl_sync(0xffffffffu, idx, 0);
    if (idx >= n_tiles_round) break;
    const int t = list_cur[idx];
    if (lane == 0) {
        atomicExch(queued + t, 0u);              // from here on a neighbour may queue me again
        __threadfence();                         // ... and I read what it wrote before it tried
    }
    __syncwarp();
    const int ty = t / tiles_x, tx = t - ty * tiles_x;
    const int r0 = ty * FT, c0 = tx * FT;
    // load: W with a one-cell rim (outside the grid: +inf), z without
    for (int i = 0; i < FT + 2; ++i) {
        const int r = r0 - 1 + i;
        const bool rin = (r >= 0 && r < ny);
        for (int jj = lane; jj < FT + 2; jj += 32) {
            const int c = c0 - 1 + jj;
            float v = INFINITY;
            if (rin && c >= 0 && c < nx) v = __ldcg(W + (int64_t)r * nx + c);
            sW[i * FSW + jj] = v;
        }
    }
    for (int i = 0; i < FT; ++i) {
        const int r = r0 + i, c = c0 + lane;
        sZ[i * FSZ + lane] = (r < ny && c < nx) ? __ldg(zz + (int64_t)r * nx + c) : INFINITY;
    }
    __syncwarp();
    unsigned rim = 0;            // bit 0 top, 1 bottom, 2 left, 3 right rim changed
    bool any = false;
#define TFB_RELAX(I, J)                                                                    \
    {                                                                                      \
        float *p = sW + (I) * FSW + (J);                                                   \
        const float cur = *p;                                                              \
        float m = fminf(fminf(p[-FSW - 1], p[-FSW]), fminf(p[-FSW + 1], p[-1]));           \
        m = fminf(m, fminf(fminf(p[1], p[FSW - 1]), fminf(p[FSW], p[FSW + 1])));           \
        const float v = fmaxf(sZ[((I) - 1) * FSZ + (J) - 1], fminf(cur, m));               \
        if (v < cur) {                                                                     \
            *p = v;                                                                        \
            ch = true;                                                                     \
            rim |= ((I) == 1 ? 1u : 0u) | ((I) == FT ? 2u : 0u) | ((J) == 1 ? 4u : 0u) |   \
                   ((J) == FT ? 8u : 0u);                                                  \
        }                                                                                  \
    }
    while (true) {
        bool ch = false;
        for (int j = 1; j <= FT; ++j) { TFB_RELAX(lane + 1, j); __syncwarp(); }      // west -> east
        for (int j = FT; j >= 1; --j) { TFB_RELAX(lane + 1, j); __syncwarp(); }      // east -> west
        for (int i = 1; i <= FT; ++i) { TFB_RELAX(i, lane + 1); __syncwarp(); }      // north -> south
        for (int i = FT; i >= 1; --i) { TFB_RELAX(i, lane + 1); __syncwarp(); }      // south -> north
        if (!__any_sync(0xffffffffu, ch)) break;
        any = true;
    }
#undef TFB_RELAX
    if (any) {
        for (int i = 0; i < FT; ++i) {
            const int r = r0 + i, c = c0 + lane;
            if (r < ny && c < nx) __stcg(W + (int64_t)r * nx + c, sW[(i + 1) * FSW + lane + 1]);
        }
        rim = __reduce_or_sync(0xffffffffu, rim);
        __threadfence();                         // my W is visible before anybody is queued
        if (lane < 8) {
            // neighbour tile `lane` (NE, E, SE, S, SW, W, NW, N) reads my rim
            const int dr = kNbrDRow[lane], dc = kNbrDCol[lane];
            const unsigned need = (dr < 0 ? 1u : (dr > 0 ? 2u : 0u)) | (dc < 0 ? 4u : (dc > 0 ? 8u : 0u));
            const int yy = ty + dr, xx = tx + dc;
            if ((rim & need) == need && yy >= 0 && yy < tiles_y && xx >= 0 && xx < tiles_x) {
                const int u = yy * tiles_x + xx;
                if (atomicExch(queued + u, 1u) == 0u) list_next[atomicAdd(n_next, 1u)] = u;
            }
        }
    }
    __syncwarp();
    }
}

__global__ void fill_list_init_kernel(int *list, unsigned *queued, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { list[i] = i; queued[i] = 1u; }
}

__global__ void __launch_bounds__(256)
fill_finish_kernel(float *dem, const float *__restrict__ W, int64_t n,
                   unsigned long long *n_raised) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool up = false;
    if (i < n) {
        const float w = W[i], z = dem[i];
        up = isfinite(w) && w > z;                              // :552-554
        if (up) dem[i] = w;
    }
    const unsigned m = __ballot_sync(0xffffffffu, up);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_raised, (unsigned long long)__popc(m));
}

// ---- power-law grid --------------------------------------------------------------------
// positive minimum and maximum of a float32 grid (bit patterns of positive floats are ordered)
__global__ void __launch_bounds__(256)
range_pos_kernel(const float *__restrict__ A, int64_t n, unsigned *out /* {min_pos_bits, max_bits} */) {
    float mn = INFINITY, mx = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float a = __ldg(A + i);
        if (a > 0.0f) { mn = fminf(mn, a); mx = fmaxf(mx, a); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(out, __float_as_uint(mn));
        atomicMax(out + 1, __float_as_uint(mx));
    }
}

__global__ void __launch_bounds__(256)
power_law_kernel(const float *__restrict__ A, int64_t n, double c, double p, int single,
                 double *__restrict__ out64, float *__restrict__ out32) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a = __ldg(A + i);
    double g;
    if (a <= 0.0f) g = 1.0;                                     // :188
    else if (single) g = (double)((float)c * (float)pow((double)a, p));   // float32 arithmetic
    else g = c * pow((double)a, p);
    if (out64) out64[i] = g;
    if (out32) out32[i] = (float)g;
}

}  // namespace tfb
using namespace tfb;

static inline int64_t align256(int64_t v) { return (v + 255) / 256 * 256; }
constexpr int FBATCH = 16;       // rounds launched between two looks at the counters

extern "C" int64_t tfb_fill_pits_scratch_bytes(int32_t ny, int32_t nx) {
    if (ny <= 0 || nx <= 0) return 0;
    const int64_t n = (int64_t)ny * nx;
    const int64_t tiles = (int64_t)((ny + FT - 1) / FT) * ((nx + FT - 1) / FT);
    // zz, W | two work lists | queued flags | counters | row flags
    return 2 * n * (int64_t)sizeof(float) + 3 * align256(tiles * 4) + 512 + align256(ny);
}

extern "C" int tfb_fill_pits(float *dem, int32_t ny, int32_t nx, float nodata, float closed_code,
                             int32_t close_nodata, void *scratch, int64_t *n_raised,
                             int32_t *n_rounds, void *stream) {
    TFB_ARG_CHECK(dem && scratch && ny > 0 && nx > 0);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = (int64_t)ny * nx;
    const int tiles_y = (ny + FT - 1) / FT, tiles_x = (nx + FT - 1) / FT;
    const int64_t tiles = (int64_t)tiles_y * tiles_x;
    TFB_ARG_CHECK(tiles < (1LL << 31));
    float *zz = (float *)scratch, *W = zz + n;
    int *list0 = (int *)(W + n);
    int *list1 = (int *)((char *)list0 + align256(tiles * 4));
    unsigned *queued = (unsigned *)((char *)list1 + align256(tiles * 4));
    // cnt[k] = length of round k's list (k = 0..FBATCH), pop[k] = tiles popped in round k
    unsigned *cnt = (unsigned *)((char *)queued + align256(tiles * 4));
    unsigned *pop = cnt + 32;
    unsigned long long *raised = (unsigned long long *)(cnt + 64);
    uint8_t *row_flag = (uint8_t *)cnt + 512;
    static_assert(FBATCH + 1 <= 32 && FBATCH % 2 == 0, "counter block");
    TFB_CUDA_CHECK(cudaFuncSetAttribute(fill_relax_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, FSMEM));
    dim3 b(32, 8), g((nx + 31) / 32, (ny + 7) / 8);
    fill_rowflag_kernel<<<ny, 256, 0, st>>>(dem, nx, nodata, row_flag);
    fill_init_kernel<<<g, b, 0, st>>>(dem, zz, W, ny, nx, nodata, closed_code, close_nodata, row_flag);
    fill_list_init_kernel<<<(unsigned)((tiles + 255) / 256), 256, 0, st>>>(list0, queued, (int)tiles);
    TFB_CUDA_CHECK(cudaGetLastError());
    int ctas_per_sm = 1;
    TFB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, fill_relax_kernel,
                                                                 32 * FWARPS, FSMEM));
    const int64_t want = (tiles + FWARPS - 1) / FWARPS;
    const int64_t cap = (int64_t)sm_count() * (ctas_per_sm > 0 ? ctas_per_sm : 1);
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    unsigned h[64];
    memset(h, 0, sizeof(h));
    h[0] = (unsigned)tiles;
    int rounds = 0;
    bool done = false;
    // every relaxation lowers at least one of finitely many values: the loop terminates;
    // the bound only guards against a broken device
    const int64_t max_rounds = 8 * ((int64_t)tiles_y + tiles_x) * FT + 1024;
    while (!done && rounds < max_rounds) {
        TFB_CUDA_CHECK(cudaMemcpyAsync(cnt, h, 64 * sizeof(unsigned), cudaMemcpyHostToDevice, st));
        for (int k = 0; k < FBATCH; ++k) {
            int *cur = (k & 1) ? list1 : list0, *nxt = (k & 1) ? list0 : list1;
            fill_relax_kernel<<<grid, 32 * FWARPS, FSMEM, st>>>(zz, W, ny, nx, tiles_y, tiles_x, cur,
                                                                nxt, queued, cnt + k, cnt + k + 1, pop + k);
        }
        TFB_CUDA_CHECK(cudaGetLastError());
        TFB_CUDA_CHECK(cudaMemcpyAsync(h, cnt, 33 * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        TFB_CUDA_CHECK(cudaStreamSynchronize(st));
        int ran = FBATCH;
        for (int k = 1; k <= FBATCH; ++k)
            if (h[k] == 0) { done = true; ran = k; break; }      // round k had nothing to do
        rounds += ran;
        const unsigned carry = h[FBATCH];
        memset(h, 0, sizeof(h));
        h[0] = carry;                                            // FBATCH is even: list0 again
    }
    if (!done) {
        set_error("tfb_fill_pits: no fixed point after %lld rounds", (long long)max_rounds);
        return TFB_ERR_NOT_CONVERGED;
    }
    TFB_CUDA_CHECK(cudaMemsetAsync(raised, 0, sizeof(*raised), st));
    fill_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dem, W, n, raised);
    TFB_CUDA_CHECK(cudaGetLastError());
    unsigned long long hr = 0;
    TFB_CUDA_CHECK(cudaMemcpyAsync(&hr, raised, sizeof(hr), cudaMemcpyDeviceToHost, st));
    TFB_CUDA_CHECK(cudaStreamSynchronize(st));
    if (n_raised) *n_raised = (int64_t)hr;
    if (n_rounds) *n_rounds = rounds;
    return TFB_OK;
}

extern "C" int tfb_area_range(const float *A, int64_t n, float *min_pos, float *max_val,
                              void *stream) {
    TFB_ARG_CHECK(A && n > 0 && min_pos && max_val);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned *d = (unsigned *)small_scratch();
    TFB_ARG_CHECK(d != nullptr);
    const unsigned init[2] = {0x7f800000u, 0u};
    cudaError_t e = cudaMemcpyAsync(d, init, sizeof(init), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        range_pos_kernel<<<sm_count() * 8, 256, 0, st>>>(A, n, d);
        e = cudaGetLastError();
    }
    unsigned h[2] = {0, 0};
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        set_error("tfb_area_range: %s", cudaGetErrorString(e));
        return TFB_ERR_CUDA;
    }
    memcpy(min_pos, &h[0], 4);
    memcpy(max_val, &h[1], 4);
    return TFB_OK;
}

extern "C" int tfb_power_law_grid(const float *A, int64_t n, double c, double p,
                                  int32_t single_precision, double *out64, float *out32,
                                  void *stream) {
    TFB_ARG_CHECK(A && n > 0 && (out64 || out32));
    power_law_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        A, n, c, p, single_precision, out64, out32);
    TFB_CUDA_CHECK(cudaGetLastError());
    return TFB_OK;
}
