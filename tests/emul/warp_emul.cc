// warp_emul.cc -- TEST INFRASTRUCTURE: the fiber scheduler behind tests/emul/cuda_runtime.h.
// One CTA at a time: `block` fibers, round-robin, switched only inside barriers.
#include "cuda_runtime.h"
#include <stdlib.h>
#include <ucontext.h>
#include <vector>

namespace emu {

struct Bar { unsigned gen = 0, arrived = 0, live = 0; };
struct Fiber { ucontext_t ctx; char *stack = nullptr; bool done = false; };

struct Cta {
	std::vector<Fiber> f;
	std::vector<Bar> wbar;                    // one per warp
	std::vector<uint64_t> x;                  // 32 exchange slots per warp
	Bar cbar;
	ucontext_t main_ctx;
	unsigned cur = 0, live = 0, block = 0, cta_id = 0, grid = 1;
	const std::function<void()> *kernel = nullptr;
};

static thread_local Cta *C = nullptr;     // one emulated device per OS thread
static unsigned long long g_syncs = 0;
static const size_t STACK = 512u << 10;

unsigned tid() { return C->cur; }
unsigned cta() { return C->cta_id; }
unsigned ncta() { return C->grid; }
unsigned nthr() { return C->block; }
uint64_t *xchg() { return &C->x[(C->cur >> 5) * 32u]; }
unsigned long long syncs() { return g_syncs; }

static void next_fiber()
{
	Cta *c = C;
	const unsigned from = c->cur;
	for (unsigned k = 1; k <= c->block; k++) {
		const unsigned l = (from + k) % c->block;
		if (!c->f[l].done) {
			if (l == from) return;
			c->cur = l;
			swapcontext(&c->f[from].ctx, &c->f[l].ctx);
			return;
		}
	}
}

static void wait_on(Bar &b)
{
	g_syncs++;
	const unsigned my = b.gen;
	if (++b.arrived >= b.live) { b.arrived = 0; b.gen++; }
	while (b.gen == my) next_fiber();
}

void barrier_warp() { wait_on(C->wbar[C->cur >> 5]); }
void barrier_cta() { wait_on(C->cbar); }

static void leave(Bar &b)
{
	// a thread that exits while others wait at a barrier no longer counts (CUDA semantics for
	// exited threads); release the barrier if everybody still alive has already arrived
	b.live--;
	if (b.live != 0 && b.arrived >= b.live) { b.arrived = 0; b.gen++; }
}

static void trampoline()
{
	Cta *c = C;
	const unsigned me = c->cur;
	(*c->kernel)();
	c->f[me].done = true;
	c->live--;
	leave(c->wbar[me >> 5]);
	leave(c->cbar);
	for (unsigned k = 1; k < c->block; k++) {
		const unsigned n = (me + k) % c->block;
		if (!c->f[n].done) { c->cur = n; setcontext(&c->f[n].ctx); }
	}
	setcontext(&c->main_ctx);
}

static void run_cta(unsigned b, unsigned grid, unsigned block, const std::function<void()> &kernel)
{
	Cta *c = new Cta();
	C = c;
	c->kernel = &kernel;
	c->block = block; c->grid = grid; c->cta_id = b; c->live = block;
	c->f.resize(block);
	c->wbar.resize((block + 31) / 32);
	c->x.assign(((block + 31) / 32) * 32, 0);
	for (unsigned w = 0; w < c->wbar.size(); w++)
		c->wbar[w].live = (w * 32 + 32 <= block) ? 32 : block - w * 32;
	c->cbar.live = block;
	for (unsigned t = 0; t < block; t++) {
		Fiber &f = c->f[t];
		f.stack = (char *)malloc(STACK);
		getcontext(&f.ctx);
		f.ctx.uc_stack.ss_sp = f.stack;
		f.ctx.uc_stack.ss_size = STACK;
		f.ctx.uc_link = &c->main_ctx;
		makecontext(&f.ctx, trampoline, 0);
	}
	c->cur = 0;
	swapcontext(&c->main_ctx, &c->f[0].ctx);
	for (unsigned t = 0; t < block; t++) free(c->f[t].stack);
	delete c;
}

void launch(unsigned grid, unsigned block, const std::function<void()> &kernel)
{
	Cta *outer = C;
	for (unsigned b = 0; b < grid; b++) run_cta(b, grid, block, kernel);
	C = outer;
}

void run_warp(const std::function<void(int)> &body)
{
	launch(1, 32, [&] { body(lane()); });
}

} // namespace emu
