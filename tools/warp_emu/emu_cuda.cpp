// emu_cuda.cpp — scheduler of the one-warp SIMT emulator (see emu_cuda.h).
#include "emu_cuda.h"
#include <vector>

namespace emu {

Warp* g_warp = nullptr;
uint32_t g_warp_index = 0;

static void lane_entry() {
    Warp* w = g_warp;
    w->body();
    w->done[w->cur] = true;
    swapcontext(&w->lanes[w->cur], &w->sched);
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
        getcontext(&w->lanes[i]);
        w->lanes[i].uc_stack.ss_sp = stacks[i];
        w->lanes[i].uc_stack.ss_size = stack_bytes;
        w->lanes[i].uc_link = &w->sched;
        makecontext(&w->lanes[i], (void (*)())lane_entry, 0);
    }
    int live = W;
    uint64_t stuck_rounds = 0;
    while (live > 0) {
        const uint64_t before = w->n_collectives * 64 + (uint64_t)(w->arrived[0] + w->arrived[1]);
        const int live_before = live;
        for (int i = 0; i < W; i++) {
            if (w->done[i]) continue;
            w->cur = i;
            swapcontext(&w->sched, &w->lanes[i]);
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
