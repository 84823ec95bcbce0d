// emu_cuda.cpp — scheduler of the one-warp SIMT emulator (see emu_cuda.h).
#include "emu_cuda.h"
#include <vector>

#ifdef EMU_FAST_SWITCH
asm(R"(
    .text
    .globl emu_switch
    .type emu_switch, @function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
    .size emu_switch, .-emu_switch
)");
#endif

namespace emu {

Warp* g_warp = nullptr;
uint32_t g_warp_index = 0;

static void lane_entry() {
    Warp* w = g_warp;
    w->body();
    w->done[w->cur] = true;
#ifdef EMU_FAST_SWITCH
    emu_switch(&w->lane_sp[w->cur], w->sched_sp);
    abort();     // a finished lane is never resumed
#else
    swapcontext(&w->lanes[w->cur], &w->sched);
#endif
}

void run_warp(const std::function<void()>& body, size_t stack_bytes) {
    Warp* w = new Warp();
    memset(w->done, 0, sizeof(w->done));
    memset(w->buf, 0, sizeof(w->buf));
    w->arrived[0] = w->arrived[1] = w->left[0] = w->left[1] = 0;
    memset(w->gen, 0, sizeof(w->gen));
    w->n_collectives = 0;
    w->body = body;
    std::vector<char*> stacks(W);
    Warp* saved = g_warp;
    g_warp = w;
    for (int i = 0; i < W; i++) {
        stacks[i] = (char*)malloc(stack_bytes);
#ifdef EMU_FAST_SWITCH
        // initial frame: six callee-saved registers (zero), the entry point as return address, a null return address for
        // the entry function itself; rsp is 8 mod 16 when lane_entry starts, as after a call
        uintptr_t top = (reinterpret_cast<uintptr_t>(stacks[i]) + stack_bytes) & ~uintptr_t(15);
        void** sp = reinterpret_cast<void**>(top);
        *--sp = nullptr;
        *--sp = reinterpret_cast<void*>(&lane_entry);
        for (int r = 0; r < 6; r++) *--sp = nullptr;
        w->lane_sp[i] = sp;
#else
        getcontext(&w->lanes[i]);
        w->lanes[i].uc_stack.ss_sp = stacks[i];
        w->lanes[i].uc_stack.ss_size = stack_bytes;
        w->lanes[i].uc_link = &w->sched;
        makecontext(&w->lanes[i], (void (*)())lane_entry, 0);
#endif
    }
    int live = W;
    uint64_t stuck_rounds = 0;
    while (live > 0) {
        const uint64_t before = w->n_collectives * 64 + (uint64_t)(w->arrived[0] + w->arrived[1]);
        const int live_before = live;
        for (int i = 0; i < W; i++) {
            if (w->done[i]) continue;
            w->cur = i;
#ifdef EMU_FAST_SWITCH
            emu_switch(&w->sched_sp, w->lane_sp[i]);
#else
            swapcontext(&w->sched, &w->lanes[i]);
#endif
            if (w->done[i]) live--;
        }
        const uint64_t after = w->n_collectives * 64 + (uint64_t)(w->arrived[0] + w->arrived[1]);
        if (after == before && live == live_before) {
            if (++stuck_rounds > 4) {
                fprintf(stderr, "warp emulator: deadlock — %d lanes wait at a collective the others never reach\n", live);
                abort();
            }
        } else stuck_rounds = 0;
    }
    for (int i = 0; i < W; i++) free(stacks[i]);
    g_warp = saved;
    delete w;
}

}  // namespace emu
