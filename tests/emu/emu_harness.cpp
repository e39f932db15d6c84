// tests/emu/emu_harness.cpp — TEST INFRASTRUCTURE: C entry points (ctypes, tests/test_emu_kernels.py) that run the kernels
// of akuaengine_b200/csrc/ on the CPU through the SIMT emulator of tests/emu/cuda_runtime.h. Built by g++ with
// -DAKUA_HOST_EMU -I tests/emu; never part of libakua_pbf.so and never loaded by the product package.
//
// emu_step() sequences the kernels exactly as stepEager() in akuaengine_b200/csrc/pbf_solver.cu does (same kernels, same
// arguments, same buffer swaps) on plain host arrays, so a whole PBF step of the CUDA source can be compared with the CPU
// port of the reference and with the golden fixtures without a GPU.
#include "../../akuaengine_b200/csrc/pbf_kernels.cuh"
#include "../../akuaengine_b200/csrc/list_build.cuh"
#include "../../akuaengine_b200/csrc/radix_sort.cuh"
#include "../../akuaengine_b200/csrc/slab_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

using namespace akua;

namespace {

template <typename... KArgs, typename... Args>
void run(void (*kernel)(KArgs...), uint32_t grid, uint32_t block, Args... args) {
    if (grid == 0) return;
    emu::launch(dim3(grid), dim3(block), [&]() { kernel(KArgs(args)...); });
}
inline uint32_t gridFor(uint64_t n, uint32_t block = 256) { return (uint32_t)((n + block - 1) / block); }
int bitsFor(uint64_t maxKey) { int b = 1; while (b < 32 && (maxKey >> b) != 0) b++; return b; }

}  // namespace

extern "C" {

// Stable LSD radix sort of (key, index) pairs: the three kernels of radix_sort.cuh driven by rsort::sort_pairs itself.
int emu_sort_pairs(const uint32_t* keysIn, uint32_t n, int keyBits, uint32_t* keysOut, uint32_t* valsOut) {
    std::vector<uint32_t> keyA(n), keyB(n), valA(n), valB(n);
    rsort::Workspace ws;
    ws.maxTiles = rsort::max_tiles_for_capacity(n);
    std::vector<uint32_t> tileHist((size_t)256 * ws.maxTiles), binTotal(256);
    ws.tileHist = tileHist.data();
    ws.binTotal = binTotal.data();
    uint32_t *ko = nullptr, *vo = nullptr;
    const int launches = rsort::sort_pairs(keysIn, keyA.data(), valA.data(), keyB.data(), valB.data(), n, keyBits, ws, nullptr, &ko, &vo, true);
    std::copy(ko, ko + n, keysOut);
    std::copy(vo, vo + n, valsOut);
    return launches;
}

}  // extern "C"
