// Cooperative-fiber emulation of one CUDA thread block on the CPU (TEST
// INFRASTRUCTURE ONLY).  Every CUDA thread is a ucontext fiber; __syncthreads
// and warp shuffles are barriers at which a fiber yields to a round-robin
// scheduler.  This lets kernel bodies written against a `Ctx` policy
// (sync / shfl_xor16) run unmodified under g++.
#pragma once
#include <ucontext.h>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <cstring>
#include <vector>

namespace emu {

struct Cta {
    int nthreads = 0;
    int current = -1;
    std::vector<ucontext_t> ctx;
    std::vector<char*> stacks;
    std::vector<char> done;
    ucontext_t sched;
    // block barrier
    int bar_count = 0, bar_gen = 0;
    // named barriers (bar.sync / bar.arrive id, count)
    int nb_count[16] = {0}, nb_gen[16] = {0};
    // tensor memory as the kernels use it: 512 32-bit columns of each thread's own lane
    std::vector<double> tmem;       // [nthreads][256]
    // per-warp shuffle state
    std::vector<double> slot;       // [nthreads]
    std::vector<int> w_count, w_gen; // [nwarps]
    std::function<void(int)> body;
};

inline Cta*& active() {
    static thread_local Cta* c = nullptr;
    return c;
}

inline void yield_now() {
    Cta* c = active();
    swapcontext(&c->ctx[c->current], &c->sched);
}

inline void trampoline() {
    Cta* c = active();
    const int tid = c->current;
    c->body(tid);
    c->done[tid] = 1;
    swapcontext(&c->ctx[tid], &c->sched);
}

inline void run_cta(int nthreads, std::function<void(int)> body, size_t stack_bytes = 512 * 1024) {
    Cta c;
    c.nthreads = nthreads;
    c.ctx.resize(nthreads);
    c.stacks.resize(nthreads);
    c.done.assign(nthreads, 0);
    c.slot.assign(nthreads, 0.0);
    c.tmem.assign((size_t)nthreads * 256, -1.2345e300);      // uninitialised columns read back as nonsense
    const int nwarps = (nthreads + 31) / 32;
    c.w_count.assign(nwarps, 0);
    c.w_gen.assign(nwarps, 0);
    c.body = std::move(body);
    Cta* prev = active();
    active() = &c;
    for (int t = 0; t < nthreads; ++t) {
        c.stacks[t] = (char*)malloc(stack_bytes);
        getcontext(&c.ctx[t]);
        c.ctx[t].uc_stack.ss_sp = c.stacks[t];
        c.ctx[t].uc_stack.ss_size = stack_bytes;
        c.ctx[t].uc_link = &c.sched;
        makecontext(&c.ctx[t], (void (*)())trampoline, 0);
    }
    int remaining = nthreads;
    while (remaining > 0) {
        for (int t = 0; t < nthreads; ++t) {
            if (c.done[t] == 1) continue;
            c.current = t;
            swapcontext(&c.sched, &c.ctx[t]);
            if (c.done[t] == 1) { c.done[t] = 2; --remaining; }
        }
        for (int t = 0; t < nthreads; ++t) if (c.done[t] == 2) c.done[t] = 1;
    }
    for (int t = 0; t < nthreads; ++t) free(c.stacks[t]);
    active() = prev;
}

// __syncthreads()
inline void sync_block() {
    Cta* c = active();
    const int gen = c->bar_gen;
    if (++c->bar_count == c->nthreads) {
        c->bar_count = 0;
        ++c->bar_gen;
    } else {
        while (c->bar_gen == gen) yield_now();
    }
}

inline void sync_warp_internal(Cta* c, int warp, int width) {
    const int gen = c->w_gen[warp];
    if (++c->w_count[warp] == width) {
        c->w_count[warp] = 0;
        ++c->w_gen[warp];
    } else {
        while (c->w_gen[warp] == gen) yield_now();
    }
}

// bar.arrive id, count / bar.sync id, count
inline void named_arrive(int id, int count) {
    Cta* c = active();
    if (++c->nb_count[id] == count) { c->nb_count[id] = 0; ++c->nb_gen[id]; }
}
inline void named_sync(int id, int count) {
    Cta* c = active();
    const int gen = c->nb_gen[id];
    if (++c->nb_count[id] == count) { c->nb_count[id] = 0; ++c->nb_gen[id]; }
    else while (c->nb_gen[id] == gen) yield_now();
}

// __shfl_xor_sync(full mask, v, mask) for whole warps
inline double shfl_xor(double v, int mask) {
    Cta* c = active();
    const int tid = c->current, warp = tid >> 5;
    const int width = (c->nthreads - warp * 32) < 32 ? (c->nthreads - warp * 32) : 32;
    c->slot[tid] = v;
    sync_warp_internal(c, warp, width);
    const double got = c->slot[tid ^ mask];
    sync_warp_internal(c, warp, width);
    return got;
}

inline void sync_warp() {
    Cta* c = active();
    const int warp = c->current >> 5;
    const int width = (c->nthreads - warp * 32) < 32 ? (c->nthreads - warp * 32) : 32;
    sync_warp_internal(c, warp, width);
}

struct EmuCtx {
    static void sync() { sync_block(); }
    static void sync_warp() { emu::sync_warp(); }
    static double shfl_xor16(double v) { return emu::shfl_xor(v, 16); }
    static double shfl_xor(double v, int mask) { return emu::shfl_xor(v, mask); }
    static void compiler_fence() {}
    // mbarrier with one arrival per phase: *bar counts completed phases; a wait on parity P returns once the
    // phase of that parity is over (the hardware's try_wait.parity); the bulk load completes when it is issued
    // bits 0-31: completed phases, 32-47: arrivals of the phase in progress, 48-63: arrivals a phase needs
    static void mbar_init(unsigned long long* bar, unsigned count = 1u) { *bar = (unsigned long long)count << 48; }
    static void mbar_arrive(unsigned long long* bar) {
        const unsigned long long need = *bar >> 48, got = ((*bar >> 32) & 0xffffull) + 1;
        if (got == need) *bar = (need << 48) | ((*bar + 1) & 0xffffffffull);
        else *bar = (need << 48) | (got << 32) | (*bar & 0xffffffffull);
    }
    static void mbar_wait(unsigned long long* bar, unsigned parity) {
        long spins = 0;
        while ((*bar & 1ull) == (unsigned long long)parity) {
            if (++spins > 2000000) {     // every other fiber has had two million turns: a deadlock of the kernel body
                fprintf(stderr, "emu: fiber %d waits forever on mbarrier %p (state %llx) for parity %u\n", active()->current,
                        (void*)bar, *bar, parity);
                abort();
            }
            emu::yield_now();
        }
    }
    static void bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
        memcpy(dst, src, bytes);
        mbar_arrive(bar);
    }
    static unsigned atomic_inc_acq_rel(unsigned* p) { return (*p)++; }
    // tensor memory (tcgen05.alloc / st / ld / dealloc): the address carries the lane quarter of the calling warp in
    // bits 16-31 -- checked, because on the device a warp cannot reach another quarter -- and the column in bits 0-15
    static uint32_t tmem_alloc(uint32_t* slot, int) { *slot = 0; sync_block(); return *slot; }
    static void tmem_free(uint32_t, int) { sync_block(); }
    static double* tmem_cols(uint32_t taddr) {
        Cta* c = active();
        const int tid = c->current;
        const uint32_t lane0 = taddr >> 16, col = taddr & 0xffffu;
        if (lane0 != (uint32_t)(((tid >> 5) & 3) * 32) || col % 4 != 0 || col + 16 > 512) {
            fprintf(stderr, "emu: thread %d addresses tensor memory lane %u column %u\n", tid, lane0, col);
            abort();
        }
        return &c->tmem[(size_t)tid * 256 + col / 2];
    }
    template <class V> static void tmem_st4(uint32_t taddr, const V (&v)[4]) {
        double* p = tmem_cols(taddr);
        for (int i = 0; i < 4; ++i) { p[2 * i] = (double)v[i].x; p[2 * i + 1] = (double)v[i].y; }
    }
    template <class V> static void tmem_ld4(uint32_t taddr, V (&v)[4]) {
        const double* p = tmem_cols(taddr);
        for (int i = 0; i < 4; ++i) { v[i].x = p[2 * i]; v[i].y = p[2 * i + 1]; }
    }
    static void tmem_wait_st() {}
    static void delay(unsigned clocks) { for (unsigned i = 0; i < clocks / 64; ++i) emu::yield_now(); }
    template <class V> static V ld_stream(const V* p) { return *p; }
};

}  // namespace emu
