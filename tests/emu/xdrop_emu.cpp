// TEST INFRASTRUCTURE.  Lane-thread emulator for bella_b200/csrc/xdrop.cuh: the device source of the X-drop kernels is
// compiled by g++ with XD_EMULATE, every lane of a warp is a fiber and the warp collectives (__shfl_sync,
// __reduce_*_sync, __all_sync, __syncwarp) go through barriers over the lanes named in the mask.  It lets the CPU suite
// run the very code the GPU runs against the oracle; it says nothing about speed.  Built by tests/emu/Makefile into
// tests/emu/_build/libxdrop_emu.so and used by tests/test_xdrop_emu.py only.
#define XD_EMULATE
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>
#if !defined(__x86_64__)
#include <ucontext.h>
#endif

#include "../../bella_b200/csrc/xdrop.cuh"

namespace {

// The 32 lanes of a warp are fibers of ONE host thread, scheduled round-robin; a lane that waits in a collective yields.
// Warps run on separate host threads.  (x86-64: a 7-instruction stack switch; elsewhere ucontext.)
#if defined(__x86_64__)
extern "C" void xd_emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl xd_emu_switch
.type xd_emu_switch,@function
xd_emu_switch:
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
.size xd_emu_switch,.-xd_emu_switch
)");
#endif

struct Bar { int count = 0; unsigned gen = 0; };

struct Warp {
	static constexpr size_t STACK = 256 << 10;
	Bar bars[32][6];                            // one barrier per (first lane, size) of a mask
	int slots[32];
	bool finished[32];
	std::function<void()> body;                 // what every lane runs
	char* stacks = nullptr;
#if defined(__x86_64__)
	void* sp[32]; void* main_sp = nullptr;
#else
	ucontext_t ctx[32], main_ctx;
#endif
	int cur = 0;
	void run();
	void yield();
};

thread_local Warp* t_warp = nullptr;

void lane_entry()
{
	Warp* w = t_warp;
	w->body();
	w->finished[w->cur] = true;
	w->yield();
	abort();                                    // a finished lane is never resumed
}

void Warp::yield()
{
#if defined(__x86_64__)
	xd_emu_switch(&sp[cur], main_sp);
#else
	swapcontext(&ctx[cur], &main_ctx);
#endif
}

void Warp::run()
{
	t_warp = this;
	stacks = (char*)aligned_alloc(64, 32 * STACK);
	for (int l = 0; l < 32; ++l) {
		finished[l] = false;
		char* top = stacks + (size_t)(l + 1) * STACK;
#if defined(__x86_64__)
		void** a = (void**)(top - 16);          // return slot: after `ret` the stack is 8 mod 16, as after a call
		a[0] = (void*)&lane_entry;
		for (int i = 1; i <= 6; ++i) a[-i] = nullptr;
		sp[l] = (void*)(a - 6);
#else
		getcontext(&ctx[l]);
		ctx[l].uc_stack.ss_sp = top - STACK; ctx[l].uc_stack.ss_size = STACK; ctx[l].uc_link = nullptr;
		makecontext(&ctx[l], lane_entry, 0);
#endif
	}
	for (bool any = true; any;) {
		any = false;
		for (int l = 0; l < 32; ++l) {
			if (finished[l]) continue;
			any = true; cur = l;
#if defined(__x86_64__)
			xd_emu_switch(&main_sp, sp[l]);
#else
			swapcontext(&main_ctx, &ctx[l]);
#endif
		}
	}
	free(stacks);
}

void bar(unsigned mask)
{
	const int n = __builtin_popcount(mask);
	Warp* w = t_warp;
	Bar& b = w->bars[__builtin_ctz(mask)][__builtin_ctz(n)];
	const unsigned g = b.gen;
	if (++b.count == n) { b.count = 0; ++b.gen; }
	else while (b.gen == g) w->yield();
}

}  // namespace

namespace xd {

int lane_id() { return t_warp->cur; }

int shfl(unsigned mask, int v, int src, int width)
{
	const int lane = t_warp->cur;
	t_warp->slots[lane] = v;
	bar(mask);
	const int r = t_warp->slots[(lane & ~(width - 1)) | (src & (width - 1))];
	bar(mask);
	return r;
}

template <class F> static int reduce(unsigned mask, int v, F f)
{
	t_warp->slots[t_warp->cur] = v;
	bar(mask);
	bool first = true; int r = 0;
	for (int l = 0; l < 32; ++l) if (mask >> l & 1) { r = first ? t_warp->slots[l] : f(r, t_warp->slots[l]); first = false; }
	bar(mask);
	return r;
}

int rmax(unsigned mask, int v) { return reduce(mask, v, [](int a, int b) { return a > b ? a : b; }); }
int rmin(unsigned mask, int v) { return reduce(mask, v, [](int a, int b) { return a < b ? a : b; }); }
bool all(unsigned mask, bool p) { return reduce(mask, p ? 1 : 0, [](int a, int b) { return a & b; }) != 0; }
void wsync(unsigned mask) { bar(mask); }
int atomic_inc(int* p) { return __atomic_fetch_add(p, 1, __ATOMIC_SEQ_CST); }

}  // namespace xd

namespace {

void run_warps(int n_warps, const std::function<void(int)>& lane_body)
{
	std::vector<Warp> warps(n_warps);
	std::vector<std::thread> th;
	for (int w = 0; w < n_warps; ++w) {
		warps[w].body = [&lane_body, w] { lane_body(w); };
		th.emplace_back([&warps, w] { warps[w].run(); });
	}
	for (auto& t : th) t.join();
}

template <int G, int T>
void run_reg(const xd::Pairs& P, const xd::Queue& Q, xd::JobResult* res, int n_warps)
{
	std::vector<std::vector<char>> rings(n_warps, std::vector<char>((32 / G) * 2 * xd::Ext<G, T>::RING));
	run_warps(n_warps, [&](int w) { xd::warp_main<G, T>(P, Q, res, rings[w].data()); });
}

template <int W>
void run_thread(const xd::Pairs& P, const xd::Queue& Q, xd::JobResult* res, int n_warps)
{
	std::vector<std::vector<int>> sc(n_warps, std::vector<int>(2 * W * 32));
	std::vector<std::vector<char>> ch(n_warps, std::vector<char>(2 * W * 32));
	run_warps(n_warps, [&](int w) { xd::thread_main<W>(P, Q, res, sc[w].data(), ch[w].data(), 32, xd::lane_id()); });
}

template <int W>
void run_thread_packed(const xd::Pairs& P, const xd::Queue& Q, xd::JobResult* res, int n_warps, const int* order)
{
	std::vector<std::vector<int>> sc(n_warps, std::vector<int>(2 * W * 32));
	run_warps(n_warps, [&](int w) { xd::thread_main_packed<W, 32>(P, Q, res, sc[w].data(), xd::lane_id(), order); });
}

template <int W>
void run_thread_two(const xd::Pairs& P, const xd::Queue& Q, xd::JobResult* res, int n_warps, const int* order)
{
	std::vector<std::vector<int>> sc(n_warps, std::vector<int>(W * 32));
	run_warps(n_warps, [&](int w) { xd::thread_main_two<W, 32>(P, Q, res, sc[w].data(), xd::lane_id(), order); });
}

void run_wide(const xd::Pairs& P, const int* list, int n_list, xd::JobResult* res, int cap, int* bad, int n_warps)
{
	std::vector<std::vector<int>> scratch(n_warps, std::vector<int>(3 * (size_t)cap));
	int next = 0;
	run_warps(n_warps, [&](int w) { xd::wide_main(P, list, n_list, &next, res, scratch[w].data(), cap, bad); });
}

}  // namespace
// G = 0: everything through the wide path.  cols == nullptr: CSC form (colptr over n_cols columns).  Returns 0, -1 for a seed outside its read, -2 for an unknown shape;
// *n_wide = extensions that left the register path.
extern "C" int xdrop_emu_align(int G, int T, uint64_t n_pairs, const uint32_t* rows, const uint32_t* cols, const uint16_t* posH,
		const uint16_t* posV, const char* seqs, const uint64_t* seq_off, uint32_t n_reads, int kmer_len, int xdrop,
		double ratiophi, double delta, int fixed_threshold, int n_warps, int32_t* out8, int* n_wide, const uint32_t* colptr, int n_cols)
{
	std::vector<char> coded(seq_off[n_reads]);                   // what k_xdrop_encode does on the device
	for (size_t i = 0; i < coded.size(); ++i) coded[i] = xd::dna5(seqs[i]);
	xd::Pairs P{rows, cols, posH, posV, coded.data(), seq_off, kmer_len, xdrop, (int)(2 * n_pairs), colptr, n_cols};
	std::vector<xd::JobResult> res(2 * n_pairs);
	std::vector<int> wide(2 * n_pairs + 1);
	int next = 0, wide_count = 0, bad = 0;
	xd::Queue Q{&next, &wide_count, wide.data(), &bad};
	int cap = 3;
	for (uint32_t r = 0; r < n_reads; ++r) cap = std::max(cap, (int)(seq_off[r + 1] - seq_off[r]) + 3);
	if (G == 0) run_wide(P, nullptr, P.n_jobs, res.data(), cap, &bad, n_warps);
	else {
		if (G == 32 && T == 1) run_reg<32, 1>(P, Q, res.data(), n_warps);
		else if (G == 32 && T == 2) run_reg<32, 2>(P, Q, res.data(), n_warps);
		else if (G == 32 && T == 4) run_reg<32, 4>(P, Q, res.data(), n_warps);
		else if (G == 16 && T == 1) run_reg<16, 1>(P, Q, res.data(), n_warps);
		else if (G == 16 && T == 2) run_reg<16, 2>(P, Q, res.data(), n_warps);
		else if (G == 8 && T == 1) run_reg<8, 1>(P, Q, res.data(), n_warps);
		else if (G == 16 && T == 4) run_reg<16, 4>(P, Q, res.data(), n_warps);
		else if (G == 8 && T == 4) run_reg<8, 4>(P, Q, res.data(), n_warps);
		else if (G == 8 && T == 8) run_reg<8, 8>(P, Q, res.data(), n_warps);
		else if (G == 1 && T == 64) run_thread<64>(P, Q, res.data(), n_warps);      // thread per extension, T = window slots
		else if (G == 1 && T == 32) run_thread<32>(P, Q, res.data(), n_warps);
		else if (G == 1 && T == 16) run_thread<16>(P, Q, res.data(), n_warps);
		else if (G == 2 || G == 3 || G == 4 || G == 5) {                             // packed thread paths; 3, 5 = longest job first
			std::vector<int> order;
			if (G == 3 || G == 5) {
				std::vector<int> est(P.n_jobs);
				for (int j = 0; j < P.n_jobs; ++j) est[j] = xd::job_estimate(P, j);
				order.resize(P.n_jobs);
				for (int j = 0; j < P.n_jobs; ++j) order[j] = j;
				std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return est[a] > est[b]; });
			}
			const int* ord = (G == 3 || G == 5) ? order.data() : nullptr;
			if (G >= 4) {                                                            // both anti-diagonals of a column in one word
				if (T == 256) run_thread_two<256>(P, Q, res.data(), n_warps, ord);
				else if (T == 128) run_thread_two<128>(P, Q, res.data(), n_warps, ord);
				else if (T == 64) run_thread_two<64>(P, Q, res.data(), n_warps, ord);
				else if (T == 32) run_thread_two<32>(P, Q, res.data(), n_warps, ord);
				else if (T == 16) run_thread_two<16>(P, Q, res.data(), n_warps, ord);
				else return -2;
			}
			else if (T == 64) run_thread_packed<64>(P, Q, res.data(), n_warps, ord);
			else if (T == 32) run_thread_packed<32>(P, Q, res.data(), n_warps, ord);
			else if (T == 16) run_thread_packed<16>(P, Q, res.data(), n_warps, ord);
			else return -2;
		}
		else return -2;
		run_wide(P, wide.data(), wide_count, res.data(), cap, &bad, n_warps);
	}
	for (uint64_t p = 0; p < n_pairs; ++p) xd::compose(P, res.data(), (int)p, ratiophi, delta, fixed_threshold, out8);
	*n_wide = wide_count;
	return bad ? -1 : 0;
}
