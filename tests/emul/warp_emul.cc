// warp_emul.cc -- TEST INFRASTRUCTURE: the fiber scheduler behind tests/emul/cuda_runtime.h
#include "cuda_runtime.h"
#include <stdlib.h>

namespace emu {

Warp *W = nullptr;
static const size_t STACK = 1u << 20;

static void next_lane()
{
	// round robin over the lanes that have not returned yet; back to main when none is left
	Warp *w = W;
	const int from = w->cur;
	for (int k = 1; k <= 32; k++) {
		const int l = (from + k) & 31;
		if (!w->done[l]) {
			if (l == from) return;
			w->cur = l;
			swapcontext(&w->ctx[from], &w->ctx[l]);
			return;
		}
	}
}

void barrier()
{
	Warp *w = W;
	w->n_sync++;
	const unsigned my = w->gen;
	if (++w->arrived == (unsigned)w->live) { w->arrived = 0; w->gen++; }
	while (w->gen == my) next_lane();
}

static void trampoline()
{
	Warp *w = W;
	const int l = w->cur;
	w->body(l);
	w->done[l] = true;
	w->live--;
	// a lane that leaves while others wait at a barrier would hang a real warp too
	if (w->arrived != 0 && w->arrived == (unsigned)w->live) { w->arrived = 0; w->gen++; }
	for (int k = 1; k < 32; k++) {
		const int n = (l + k) & 31;
		if (!w->done[n]) { w->cur = n; setcontext(&w->ctx[n]); }
	}
	setcontext(&w->main_ctx);
}

void run_warp(const std::function<void(int)> &body)
{
	Warp *w = new Warp();
	Warp *outer = W;
	W = w;
	w->body = body;
	w->live = 32;
	for (int l = 0; l < 32; l++) {
		w->done[l] = false;
		w->stack[l] = (char *)malloc(STACK);
		getcontext(&w->ctx[l]);
		w->ctx[l].uc_stack.ss_sp = w->stack[l];
		w->ctx[l].uc_stack.ss_size = STACK;
		w->ctx[l].uc_link = &w->main_ctx;
		makecontext(&w->ctx[l], trampoline, 0);
	}
	w->cur = 0;
	swapcontext(&w->main_ctx, &w->ctx[0]);
	for (int l = 0; l < 32; l++) free(w->stack[l]);
	W = outer;
	delete w;
}

} // namespace emu
