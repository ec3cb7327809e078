// Host-side check of the tile walkers the routing kernels use (tfb_pk.cuh: tile_of<PEER>,
// TileWalk<PEER> are __host__ __device__): for many grid shapes, grid sizes and both orders,
// the division-free walker must visit exactly the tiles tile_of() names, every tile of the
// grid must be visited exactly once over all CTAs, and with the peer-halo order the first and
// last tile rows must come before any interior tile of the same CTA.  Compiled and run by
// tests/test_host_cpu.py with nvcc as plain host code (no GPU needed).
#include <cstdio>
#include <vector>

#include "../../topoflow36_b200/csrc/tfb_pk.cuh"

using namespace tfb;

template <bool PEER>
static int check(int tiles_x, int tiles_y, int grid, int reverse) {
    PkConsts k{};
    k.tiles_x = tiles_x; k.tiles_y = tiles_y; k.n_tiles = tiles_x * tiles_y;
    k.n_boundary = tiles_x * (tiles_y > 1 ? 2 : 1);
    k.reverse = reverse;
    if (grid > k.n_tiles) grid = k.n_tiles;
    std::vector<int> seen((size_t)k.n_tiles, 0);
    for (int b = 0; b < grid; ++b) {
        TileWalk<PEER> w;
        w.init(k, b);
        bool interior_seen = false;
        for (int t = b; t < k.n_tiles; t += grid) {
            int ty, tx;
            tile_of<PEER>(k, t, ty, tx);
            if (ty != w.ty || tx != w.tx) {
                std::printf("walker (%d,%d) != tile_of (%d,%d): tiles %dx%d grid %d cta %d t %d peer %d rev %d\n",
                            w.ty, w.tx, ty, tx, tiles_x, tiles_y, grid, b, t, (int)PEER, reverse);
                return 1;
            }
            if (ty < 0 || ty >= tiles_y || tx < 0 || tx >= tiles_x) { std::printf("out of range\n"); return 1; }
            if (PEER) {
                const bool boundary = (ty == 0 || ty == tiles_y - 1);
                if (boundary && interior_seen) { std::printf("boundary tile after an interior one\n"); return 1; }
                if (!boundary) interior_seen = true;
            }
            seen[(size_t)ty * tiles_x + tx] += 1;
            w.advance(k, grid);
        }
    }
    for (int i = 0; i < k.n_tiles; ++i)
        if (seen[(size_t)i] != 1) { std::printf("tile %d visited %d times\n", i, seen[(size_t)i]); return 1; }
    return 0;
}

int main() {
    const int shapes[][2] = {{1, 1}, {1, 7}, {7, 1}, {2, 2}, {3, 2}, {64, 171}, {64, 274}, {1024, 342},
                             {5, 3}, {149, 2}, {148, 3}, {37, 41}, {512, 1093}, {1, 148}, {200, 1}};
    const int grids[] = {1, 2, 7, 128, 147, 148};
    int n = 0;
    for (auto &s : shapes)
        for (int g : grids) {
            if (check<false>(s[0], s[1], g, 0)) return 1;
            if (check<false>(s[0], s[1], g, 1)) return 1;
            if (check<true>(s[0], s[1], g, 0)) return 1;
            n += 3;
        }
    std::printf("ok %d cases\n", n);
    return 0;
}
