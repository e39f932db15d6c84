// list_build_host.cu — TEST INFRASTRUCTURE (not part of the product library): runs the __host__ __device__ per-particle
// list-build function of akuaengine_b200/csrc/list_build.cuh on the CPU, next to an independent one-candidate-at-a-time
// scan written here, so that the two-phase mask variant can be checked bit-for-bit without a GPU
// (tests/test_list_build_host.py). Compiled with nvcc (host code only is executed).
//
// The scan mirrors what kernel_find_neighbours does (src/CUDA/NeighbourSearchCUDA.cu:72-130: cells dx -> dy -> dz, bucket
// order, strict d2 < h*h, self skipped, capped at maxNeighbours) on the LINEAR_CELL structures of this repo (cell ranges
// instead of the hash table).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

#include "../../akuaengine_b200/csrc/list_build.cuh"

using namespace akua;

namespace {

// same layout rule as layoutGrid() in pbf_solver.cu: box + two-cell margin
GridParams make_grid(const float* bmin, const float* bmax, float h) {
    GridParams G{};
    G.cellSize = h; G.lookupCellSize = h; G.tableSize = 0;
    int lo[3], dim[3];
    for (int a = 0; a < 3; a++) {
        lo[a] = (int)std::floor(bmin[a] / h) - 2;
        int hi = (int)std::floor(bmax[a] / h) + 2;
        dim[a] = hi - lo[a] + 1;
    }
    G.gridMin = make_int3(lo[0], lo[1], lo[2]);
    G.gridDim = make_int3(dim[0], dim[1], dim[2]);
    return G;
}

uint32_t scan_reference(uint32_t i, const std::vector<float4>& xs, const std::vector<uint2>& cellRange, uint32_t maxN,
                        const GridParams& G, float h, uint32_t* out) {
    const float4 xi = xs[i];
    const float h2 = h * h;
    const int3 c = cell_of(xi.x, xi.y, xi.z, G.lookupCellSize);
    const int cx = clampi(c.x - G.gridMin.x, 0, G.gridDim.x - 1);
    const int cy = clampi(c.y - G.gridMin.y, 0, G.gridDim.y - 1);
    const int cz = clampi(c.z - G.gridMin.z, 0, G.gridDim.z - 1);
    uint32_t count = 0;
    for (int dx = -1; dx <= 1; dx++)
        for (int dy = -1; dy <= 1; dy++)
            for (int dz = -1; dz <= 1; dz++) {
                const int X = cx + dx, Y = cy + dy, Z = cz + dz;
                if (X < 0 || X >= G.gridDim.x || Y < 0 || Y >= G.gridDim.y || Z < 0 || Z >= G.gridDim.z) continue;
                const uint2 r = cellRange[((size_t)X * G.gridDim.y + Y) * G.gridDim.z + Z];
                for (uint32_t cand = r.x; cand < r.y; cand++) {
                    if (count >= maxN) return count;
                    if (cand == i) continue;
                    const float4 xj = xs[cand];
                    if (dist2(xi.x - xj.x, xi.y - xj.y, xi.z - xj.z) < h2) out[count++] = cand;
                }
            }
    return count;
}

}  // namespace

// xyz: n positions (3 floats each) in any order. Sorts them by LINEAR_CELL key (stable), builds the cell ranges and then
// the neighbour lists of every particle with `variant` (0 = independent scan, 1 = build_list_mask<4>, 2 = <8>).
// order[k] = input index of sorted slot k; lists come back row-major, maxN entries per particle, unused entries
// 0xffffffff; cnt[k] = neighbour count of sorted slot k. Returns 0, or a negative number on bad arguments.
extern "C" int akua_test_list_build_host(const float* xyz, uint32_t n, const float* bmin, const float* bmax, float h,
                                         uint32_t maxN, int variant, uint32_t* order, uint32_t* listRowMajor,
                                         uint32_t* cnt) {
    if (!xyz || !order || !listRowMajor || !cnt || n == 0 || maxN == 0 || variant < 0 || variant > 2) return -1;
    const GridParams G = make_grid(bmin, bmax, h);
    const size_t cells = (size_t)G.gridDim.x * G.gridDim.y * G.gridDim.z;
    std::vector<uint32_t> keys(n);
    for (uint32_t i = 0; i < n; i++)
        keys[i] = linear_key(cell_of(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], G.cellSize), G);
    std::vector<uint32_t> perm(n);
    std::iota(perm.begin(), perm.end(), 0u);
    std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
    std::vector<float4> xs(n);
    std::vector<uint2> cellRange(cells, make_uint2(0, 0));
    for (uint32_t k = 0; k < n; k++) {
        const uint32_t src = perm[k];
        order[k] = src;
        xs[k] = make_float4(xyz[3 * src], xyz[3 * src + 1], xyz[3 * src + 2], 1.0f);
        const uint32_t key = keys[src];
        if (k == 0 || keys[perm[k - 1]] != key) cellRange[key].x = k;
        if (k == n - 1 || keys[perm[k + 1]] != key) cellRange[key].y = k + 1;
    }
    std::fill(listRowMajor, listRowMajor + (size_t)n * maxN, 0xffffffffu);
    if (variant == 0) {
        for (uint32_t i = 0; i < n; i++) cnt[i] = scan_reference(i, xs, cellRange, maxN, G, h, listRowMajor + (size_t)i * maxN);
        return 0;
    }
    const uint32_t stride = (n + 31u) / 32u * 32u;
    const uint32_t groups = (maxN + 3u) / 4u;
    std::vector<uint32_t> ell((size_t)stride * groups * 4, 0xdeadbeefu);
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t c = variant == 1 ? build_list_mask<4>(i, xs.data(), cellRange.data(), stride, maxN, ell.data(), G, h)
                                        : build_list_mask<8>(i, xs.data(), cellRange.data(), stride, maxN, ell.data(), G, h);
        finish_list(i, c, stride, ell.data(), cnt);
    }
    for (uint32_t i = 0; i < n; i++) {
        if (cnt[i] > maxN) return -2;
        for (uint32_t k = 0; k < cnt[i]; k++) listRowMajor[(size_t)i * maxN + k] = ell[list_slot(i, k, stride)];
        // the padding of the last group must hold the particle's own index, and nothing may be written past it
        for (uint32_t k = cnt[i]; k < groups * 4; k++) {
            const uint32_t v = ell[list_slot(i, k, stride)];
            const bool pad = k < ((cnt[i] + 3u) & ~3u);
            if (pad ? v != i : v != 0xdeadbeefu) return -3;
        }
    }
    return 0;
}
