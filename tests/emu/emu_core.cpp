// tests/emu/emu_core.cpp — TEST INFRASTRUCTURE: the fiber scheduler behind tests/emu/cuda_runtime.h (see there).
#include "cuda_runtime.h"

#include <sched.h>
thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;
void emu_pause() { sched_yield(); }

namespace emu {

thread_local Cta* g_cta = nullptr;
thread_local long long g_clock = 0;

namespace {
constexpr size_t kStackBytes = 64 * 1024;
thread_local std::vector<char> g_stacks;
thread_local uint64_t g_progress = 0;
thread_local int g_live = 0, g_atBarrier = 0;

// One fiber per thread slot, created once per launch and re-used for every CTA of the grid: after the kernel body returns
// the fiber parks in the scheduler and runs the body again when it is resumed for the next CTA.
void fiber_main() {
    Cta* c = g_cta;
    for (;;) {
        c->body();
        Thread& th = c->threads[c->current];
        th.done = true;
        yield();
    }
}

void try_release_barrier(Cta& c) {
    if (g_live > 0 && g_atBarrier == g_live) {
        for (Thread& t : c.threads) t.atBarrier = false;
        g_atBarrier = 0;
        c.barrierGen++;
        g_progress++;
    }
}
}  // namespace

// EMU_USE_SWAPCONTEXT: every switch through swapcontext (slower: two signal-mask syscalls each) — what AddressSanitizer's
// interceptors understand; used by the sanitizer run of tests/test_emu_kernels.py.
void yield() {
    Cta& c = *g_cta;
#ifdef EMU_USE_SWAPCONTEXT
    swapcontext(&c.threads[c.current].ctx, &c.sched);
#else
    if (_setjmp(c.threads[c.current].jb) == 0) _longjmp(c.schedJb, 1);
#endif
}

uint32_t live_mask(int warp) {
    Cta& c = *g_cta;
    uint32_t m = 0;
    const int base = warp * 32, n = (int)c.threads.size();
    for (int l = 0; l < 32 && base + l < n; l++)
        if (!c.threads[base + l].done) m |= 1u << l;
    return m;
}

void warp_exchange(uint32_t mask, uint64_t mine, uint64_t out[32], uint32_t* participants) {
    Cta& c = *g_cta;
    const int t = c.current, w = t >> 5, lane = t & 31;
    Thread& th = c.threads[t];
    Warp& W = c.warps[w];
    const uint64_t g = th.gen;
    const int b = (int)(g & 1);
    if (W.bufGen[b] != g) { W.bufGen[b] = g; W.arrived[b] = 0; }
    W.val[b][lane] = mine;
    W.arrived[b] |= 1u << lane;
    g_progress++;
    for (;;) {
        const uint32_t live = live_mask(w);
        if ((mask & live) != live) {
            fprintf(stderr, "emu: warp collective with mask %08x but live lanes %08x (sub-warp masks are not modelled)\n", mask, live);
            abort();
        }
        if ((W.arrived[b] & live) == live) break;
        yield();
    }
    std::memcpy(out, W.val[b], sizeof(W.val[b]));
    *participants = W.arrived[b];
    th.gen = g + 1;
}

void launch(dim3 grid, dim3 block, const std::function<void()>& body) {
    if (block.y != 1 || block.z != 1 || grid.y != 1 || grid.z != 1) { fprintf(stderr, "emu: only 1-D launches are modelled\n"); abort(); }
    const int n = (int)block.x;
    if (g_stacks.size() < (size_t)n * kStackBytes) g_stacks.resize((size_t)n * kStackBytes);
    Cta c;
    c.body = body;
    c.threads.resize(n);
    c.warps.resize((n + 31) / 32);
    Cta* saved = g_cta;
    g_cta = &c;
    ::gridDim = grid;
    ::blockDim = block;
    for (int t = 0; t < n; t++) {
        Thread& th = c.threads[t];
        th.started = false;
        getcontext(&th.ctx);
        th.ctx.uc_stack.ss_sp = g_stacks.data() + (size_t)t * kStackBytes;
        th.ctx.uc_stack.ss_size = kStackBytes;
        th.ctx.uc_link = &c.sched;
        makecontext(&th.ctx, fiber_main, 0);
    }
    for (uint32_t bx = 0; bx < grid.x; bx++) {
        ::blockIdx = uint3{bx, 0, 0};
        for (Thread& th : c.threads) { th.done = false; th.gen = 0; th.barrierGen = 0; th.atBarrier = false; }
        for (Warp& w : c.warps) w = Warp{};
        c.barrierGen = 0;
        g_live = n;
        g_atBarrier = 0;
        int remaining = n;
        while (remaining > 0) {
            const uint64_t before = g_progress;
            for (int t = 0; t < n; t++) {
                Thread& th = c.threads[t];
                if (th.done) continue;
                c.current = t;
                ::threadIdx = uint3{(uint32_t)t, 0, 0};
#ifdef EMU_USE_SWAPCONTEXT
                swapcontext(&c.sched, &th.ctx);
#else
                if (_setjmp(c.schedJb) == 0) {
                    if (!th.started) { th.started = true; setcontext(&th.ctx); }
                    else _longjmp(th.jb, 1);
                }
#endif
                if (th.done) {
                    remaining--;
                    g_live--;
                    g_progress++;
                    try_release_barrier(c);   // exited threads count as arrived
                }
            }
            if (g_progress == before && remaining > 0) {
                fprintf(stderr, "emu: deadlock in block %u (%d threads parked, none can make progress)\n", bx, remaining);
                abort();
            }
        }
    }
    g_cta = saved;
}

}  // namespace emu

void __syncthreads() {
    using namespace emu;
    Cta& c = *g_cta;
    Thread& th = c.threads[c.current];
    const uint64_t g = c.barrierGen;
    th.atBarrier = true;
    g_atBarrier++;
    g_progress++;
    try_release_barrier(c);
    while (c.barrierGen == g) yield();
}
