// l1_gather_probe.cu — micro-benchmark: what does ONE warp-wide gather cost the L1 on sm_100a as a function of the address
// pattern (distinct 128-byte lines, distinct 32-byte sectors, 16-byte slot collisions inside a quarter-warp)?
//
// Why: every neighbour sweep of the PBF step is bound by l1tex data-stage wavefronts, a real sweep's
// gather touches ~12 lines / ~20 sectors with ~27 active lanes (tests/gather_locality_study.py, computed on the CPU), and ncu
// reports ~9.8 wavefronts for it — but three cost models fit that one number (one wavefront per line, one per two sectors,
// 2.45x slot collisions per quarter-warp). They predict different things for alternative layouts, so this probe measures
// them apart. Build and run on a B200 (tools/r02_first_call.sh):
//     nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_build/l1_gather_probe tools/l1_gather_probe.cu
//     tools/_build/l1_gather_probe            # prints one JSON line per (pattern, element size)
// Cycles per warp-gather = total time x nominal SM clock / (gather instructions issued per SM). Data is L1-resident (the warps of a CTA re-walk one 8 KB window).
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int kWindow = 512;          // float4 elements per warp window (8 KB): 54 neighbour cells of 8 particles, rounded up
constexpr int kWarpsPerCta = 8;
constexpr int kIters = 2048;

struct Pattern { const char* name; int lines, sectors; uint32_t off[32]; uint32_t activeMask; };

template <typename T> struct Ld;
template <> struct Ld<float>  { static __device__ float get(const float4* p)  { return __ldg(reinterpret_cast<const float*>(p)); } };
template <> struct Ld<float2> { static __device__ float get(const float4* p)  { float2 v = __ldg(reinterpret_cast<const float2*>(p)); return v.x + v.y; } };
template <> struct Ld<float4> { static __device__ float get(const float4* p)  { float4 v = __ldg(p); return v.x + v.y + v.z + v.w; } };
struct F8 { float v[8]; };
template <> struct Ld<F8>     { static __device__ float get(const float4* p)  {
    float a, b, c, d, e, f, g, h;   // one 256-bit load (sm_100+); the address must be 32-byte aligned: callers use even offsets
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d), "=f"(e), "=f"(f), "=f"(g), "=f"(h) : "l"(p));
    return a + b + c + d + e + f + g + h; } };

// Each lane gathers element (window base of its warp) + ((off + rot) mod kWindow): `rot` advances by a multiple of 8 elements
// per iteration, so line / sector / slot relations between the lanes are those of the pattern in every iteration.
template <typename T>
__global__ void __launch_bounds__(32 * kWarpsPerCta) k_probe(const float4* __restrict__ data, const uint32_t* __restrict__ offs,
                                                            uint32_t activeMask, float* __restrict__ sink, int iters) {
    const int lane = threadIdx.x & 31;
    const float4* win = data + (size_t)blockIdx.x * kWindow;   // one 8 KB window per CTA: L1-resident whatever the occupancy
    const uint32_t off = offs[lane];
    float acc = 0.f;
    if ((activeMask >> lane) & 1u) {
        uint32_t rot = 0;
        for (int it = 0; it < iters; it += 4) {
            // four independent gathers in flight, like neighbour_sweep
            const float a = Ld<T>::get(win + ((off + rot) & (kWindow - 1)));
            const float b = Ld<T>::get(win + ((off + rot + 40) & (kWindow - 1)));
            const float c = Ld<T>::get(win + ((off + rot + 80) & (kWindow - 1)));
            const float d = Ld<T>::get(win + ((off + rot + 120) & (kWindow - 1)));
            acc += (a + b) + (c + d);
            rot += 168;   // multiple of 8: keeps alignment classes
        }
    }
    if (acc == 123.456f) sink[0] = acc;   // never true for the zero-filled data: keeps the loads alive
}

static std::vector<Pattern> patterns() {
    std::vector<Pattern> P;
    auto add = [&](const char* name, int lines, int sectors, auto f, uint32_t mask = 0xffffffffu) {
        Pattern p{}; p.name = name; p.lines = lines; p.sectors = sectors; p.activeMask = mask;
        for (int l = 0; l < 32; l++) p.off[l] = f(l);
        P.push_back(p);
    };
    add("coalesced (4 lines, 16 sectors)", 4, 16, [](int l) { return (uint32_t)l; });
    add("broadcast (1 address)", 1, 1, [](int) { return 0u; });
    add("quarter-warps broadcast (4 addresses in 4 lines)", 4, 4, [](int l) { return (uint32_t)(l / 8) * 40u; });
    add("32 lines, same 16-B slot in every line", 32, 32, [](int l) { return (uint32_t)l * 8u; });
    add("32 lines, slot = lane % 8 (no slot collision inside a quarter-warp)", 32, 32, [](int l) { return (uint32_t)l * 8u + (l % 8); });
    add("16 lines, lane pairs share a sector", 16, 16, [](int l) { return (uint32_t)(l / 2) * 8u + (l % 2); });
    add("16 lines, lane pairs in two sectors of a line", 16, 32, [](int l) { return (uint32_t)(l / 2) * 8u + (l % 2) * 2u; });
    add("8 lines, 4 lanes per line in 2 sectors", 8, 16, [](int l) { return (uint32_t)(l / 4) * 8u + (l % 4); });
    add("8 lines, 4 lanes per line in 4 sectors", 8, 32, [](int l) { return (uint32_t)(l / 4) * 8u + (l % 4) * 2u; });
    add("4 lines, quarter-warp per line, 8 slots (= coalesced, lines apart)", 4, 16, [](int l) { return (uint32_t)(l / 8) * 40u + (l % 8); });
    add("12 lines / 20 sectors, slots random (like a real sweep gather), 27 active lanes", 12, 20, [](int l) {
        // 12 lines; lanes 0..19 open 20 distinct sectors, lanes 20..26 re-use sectors; pseudo-random slot inside the sector
        static const int line_of[27] = {0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 9, 10, 11, 0, 2, 4, 6, 8, 10, 11};
        static const int sect_of[27] = {0, 1, 0, 2, 1, 3, 0, 1, 2, 3, 0, 2, 1, 3, 0, 1, 2, 0, 3, 1, 0, 1, 2, 1, 2, 3, 1};
        if (l >= 27) return 0u;
        return (uint32_t)(line_of[l] * 8 + sect_of[l] * 2 + ((l * 7) & 1)); }, 0x07ffffffu);
    add("27 lanes, 27 lines, random slot", 27, 27, [](int l) { return l >= 27 ? 0u : (uint32_t)(l * 8 + ((l * 5 + 3) & 7)); }, 0x07ffffffu);
    add("27 lanes coalesced", 4, 14, [](int l) { return (uint32_t)l; }, 0x07ffffffu);
    return P;
}

template <typename T>
static void run(const char* tname, int bytes, const std::vector<Pattern>& P, const float4* data, uint32_t* dOffs, float* sink,
                int ctas, double smClockHz, int sms) {
    for (const Pattern& p : P) {
        uint32_t offs[32];
        for (int l = 0; l < 32; l++) offs[l] = bytes == 32 ? (p.off[l] & ~1u) : p.off[l];   // 256-bit loads need even elements
        CK(cudaMemcpy(dOffs, offs, sizeof(offs), cudaMemcpyHostToDevice));
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int rep = 0; rep < 2; rep++) {   // first repetition warms the L1 / clocks
            CK(cudaEventRecord(e0));
            k_probe<T><<<ctas, 32 * kWarpsPerCta>>>(data, dOffs, p.activeMask, sink, kIters);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
        }
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double gathersPerSm = (double)ctas * kWarpsPerCta * kIters / sms;
        const double cyc = ms * 1e-3 * smClockHz / gathersPerSm;
        printf("{\"elem\": \"%s\", \"bytes\": %d, \"pattern\": \"%s\", \"lines\": %d, \"sectors\": %d, \"ms\": %.4f, "
               "\"sm_cycles_per_warp_gather\": %.2f}\n", tname, bytes, p.name, p.lines, p.sectors, ms, cyc);
        fflush(stdout);
        CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
    }
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int clkKHz = 0;
    CK(cudaDeviceGetAttribute(&clkKHz, cudaDevAttrClockRate, 0));
    const int ctas = sms * 8;                       // one full wave of 256-thread CTAs
    const size_t elems = (size_t)ctas * kWindow;
    float4* data; uint32_t* dOffs; float* sink;
    CK(cudaMalloc(&data, elems * sizeof(float4)));
    CK(cudaMemset(data, 0, elems * sizeof(float4)));
    CK(cudaMalloc(&dOffs, 32 * sizeof(uint32_t)));
    CK(cudaMalloc(&sink, sizeof(float)));
    printf("{\"device\": \"%s\", \"sms\": %d, \"sm_clock_khz_nominal\": %d, \"note\": \"cycles use the nominal clock; compare patterns, "
           "not absolutes\"}\n", prop.name, sms, clkKHz);
    const std::vector<Pattern> P = patterns();
    const double hz = clkKHz * 1e3;
    run<float4>("float4", 16, P, data, dOffs, sink, ctas, hz, sms);
    run<float>("float", 4, P, data, dOffs, sink, ctas, hz, sms);
    run<float2>("float2", 8, P, data, dOffs, sink, ctas, hz, sms);
    run<F8>("8 x float (LDG.256)", 32, P, data, dOffs, sink, ctas, hz, sms);
    return 0;
}
