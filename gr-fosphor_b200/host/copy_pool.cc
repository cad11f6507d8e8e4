/*
 * copy_pool.cc - see copy_pool.h.
 *
 * Job protocol.  start() publishes the job (fields, then open_, then a
 * generation bump); workers and the calling thread claim fixed-size items
 * with one fetch_add each, in address order, so pieces complete in order;
 * the caller's wait on the LAST piece closes the job (open_ = false) and
 * waits until no worker is inside run_item() any more - only then may the
 * next start() rewrite the fields.  A worker announces itself (active_++)
 * BEFORE it looks at open_, the closer clears open_ BEFORE it looks at
 * active_ (both seq_cst): either the closer sees the worker or the worker
 * sees the job closed.
 */
#include "copy_pool.h"

#include <chrono>
#include <cstdlib>
#include <cstring>

#include <sched.h>
#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define CPU_RELAX() _mm_pause()
#else
#define CPU_RELAX() ((void)0)
#endif

namespace fosphor_b200 {

namespace {
#if defined(__x86_64__)
/* non-temporal copy: the destination (page-locked staging memory the DMA engine reads next) is
 * written around the caches, which saves the read-for-ownership of every destination line - one of
 * the four trips a staged byte makes through host DRAM */
__attribute__((target("avx2"))) void copy_stream_avx2(char *d, const char *s, size_t n)
{
	size_t head = (32 - (reinterpret_cast<uintptr_t>(d) & 31)) & 31;
	if (head > n)
		head = n;
	memcpy(d, s, head);
	d += head; s += head; n -= head;
	for (; n >= 128; n -= 128, d += 128, s += 128) {
		const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s));
		const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + 32));
		const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + 64));
		const __m256i e = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + 96));
		_mm256_stream_si256(reinterpret_cast<__m256i *>(d), a);
		_mm256_stream_si256(reinterpret_cast<__m256i *>(d + 32), b);
		_mm256_stream_si256(reinterpret_cast<__m256i *>(d + 64), c);
		_mm256_stream_si256(reinterpret_cast<__m256i *>(d + 96), e);
	}
	_mm_sfence();
	memcpy(d, s, n);
}
#endif
constexpr size_t MIN_ITEM = 64 * 1024;
/* how long an idle worker keeps polling before it sleeps: a streaming caller comes back every
 * 150-400 us (8 MiB per call at PCIe speed, a read-back in between); a sleeping worker costs
 * 50-100 us to wake, which was most of the staging time when this was 200 us (2.9 -> 5.1 Gsamples/s
 * staged).  Longer windows (4 ms) measured no better and burn a core per worker in a sink that
 * renders 30 frames a second */
constexpr auto SPIN_FOR = std::chrono::microseconds(1500);
} /* namespace */

int copy_pool::default_threads()
{
	if (const char *v = getenv("FOSPHOR_B200_COPY_THREADS")) {
		const int t = atoi(v);
		if (t >= 1)
			return t > 32 ? 32 : t;
	}
	int cpus = 0;
	cpu_set_t set;
	if (sched_getaffinity(0, sizeof(set), &set) == 0)
		cpus = CPU_COUNT(&set);
	if (cpus < 1)
		cpus = (int)std::thread::hardware_concurrency();
	if (cpus < 1)
		cpus = 2;
	/* one process per GPU (torchrun exports LOCAL_WORLD_SIZE): share the host cores */
	int local_world = 1;
	if (const char *v = getenv("LOCAL_WORLD_SIZE"))
		local_world = atoi(v) > 0 ? atoi(v) : 1;
	/* ... unless somebody already did: an affinity mask narrower than the machine is this process's
	 * own share (bench.py binds each rank to its cores) */
	const int online = (int)std::thread::hardware_concurrency();
	if (online > 0 && cpus < online)
		local_world = 1;
	int t = cpus / local_world / 2;      /* the caller copies too; leave cores for the producer side */
	if (t < 1) t = 1;
	/* measured on the GPU box (tools/e2e_probe.py, 8 MiB calls, 53 GB/s PCIe): 4 workers + the caller
	 * keep up with the DMA engine (5 x 12 GB/s); 6, 8, 12 are each a little SLOWER (5.4 / 5.2 / 5.15 /
	 * 4.98 Gsamples/s): the copy only has to keep pace, and more threads fight the DMA reads for
	 * host memory bandwidth */
	if (t > 4) t = 4;
	return t;
}

copy_pool::copy_pool(int n_threads)
{
	for (auto &d : piece_done_)
		d.store(0, std::memory_order_relaxed);
	if (n_threads <= 0)
		n_threads = default_threads();
	workers_.reserve((size_t)n_threads);
	for (int i = 0; i < n_threads; i++)
		workers_.emplace_back([this] { worker(); });
}

copy_pool::~copy_pool()
{
	stop_.store(true, std::memory_order_seq_cst);
	generation_.fetch_add(1, std::memory_order_seq_cst);
	{
		std::lock_guard<std::mutex> lk(mtx_);
		cv_.notify_all();
	}
	for (auto &t : workers_)
		t.join();
}

bool copy_pool::run_item()
{
	{
		if (next_item_.load(std::memory_order_relaxed) >= n_items_)
			return false;
		const int it = next_item_.fetch_add(1, std::memory_order_acq_rel);
		if (it >= n_items_)
			return false;
		const int piece = it / items_per_piece_, k = it - piece * items_per_piece_;
		const size_t p0 = piece_offset(piece), psz = piece_size(piece);
		const size_t off = (size_t)k * item_bytes_;
		if (off < psz) {
			const size_t n = psz - off < item_bytes_ ? psz - off : item_bytes_;
#if defined(__x86_64__)
			if (nt_)
				copy_stream_avx2(dst_ + p0 + off, src_ + p0 + off, n);
			else
#endif
				memcpy(dst_ + p0 + off, src_ + p0 + off, n);
		}
		piece_done_[piece].fetch_add(1, std::memory_order_release);
	}
	return true;
}

void copy_pool::worker()
{
	unsigned seen = generation_.load(std::memory_order_acquire);
	while (!stop_.load(std::memory_order_acquire)) {
		/* wait for the next job: spin first, then sleep */
		const auto t0 = std::chrono::steady_clock::now();
		unsigned spins = 0;
		while (generation_.load(std::memory_order_acquire) == seen) {
			CPU_RELAX();
			if ((++spins & 255u) == 0 && std::chrono::steady_clock::now() - t0 > SPIN_FOR) {
				std::unique_lock<std::mutex> lk(mtx_);
				sleepers_.fetch_add(1, std::memory_order_seq_cst);
				cv_.wait(lk, [&] { return generation_.load(std::memory_order_acquire) != seen; });
				sleepers_.fetch_sub(1, std::memory_order_seq_cst);
				break;
			}
		}
		seen = generation_.load(std::memory_order_acquire);
		if (stop_.load(std::memory_order_acquire))
			break;
		active_.fetch_add(1, std::memory_order_seq_cst);
		if (open_.load(std::memory_order_seq_cst))
			while (run_item()) { }
		active_.fetch_sub(1, std::memory_order_seq_cst);
	}
}

void copy_pool::start(void *dst, const void *src, size_t bytes, int pieces, bool non_temporal)
{
#if defined(__x86_64__)
	nt_ = non_temporal && __builtin_cpu_supports("avx2");
#else
	nt_ = false;
#endif
	if (pieces < 1) pieces = 1;
	if (pieces > MAX_PIECES) pieces = MAX_PIECES;
	dst_ = static_cast<char *>(dst);
	src_ = static_cast<const char *>(src);
	bytes_ = bytes;
	pieces_ = pieces;
	/* pieces are multiples of 4 KiB so that every piece but the last has the same size */
	piece_bytes_ = ((bytes + (size_t)pieces - 1) / (size_t)pieces + 4095) & ~(size_t)4095;
	if (piece_bytes_ == 0)
		piece_bytes_ = 4096;
	while (pieces_ > 1 && (size_t)(pieces_ - 1) * piece_bytes_ >= bytes)
		pieces_--;
	const size_t parties = workers_.size() + 1;
	item_bytes_ = ((piece_bytes_ + parties - 1) / parties + 4095) & ~(size_t)4095;
	if (item_bytes_ < MIN_ITEM)
		item_bytes_ = MIN_ITEM;
	items_per_piece_ = (int)((piece_bytes_ + item_bytes_ - 1) / item_bytes_);
	n_items_ = items_per_piece_ * pieces_;
	for (int p = 0; p < pieces_; p++)
		piece_done_[p].store(0, std::memory_order_relaxed);
	next_item_.store(0, std::memory_order_relaxed);
	open_.store(true, std::memory_order_seq_cst);
	generation_.fetch_add(1, std::memory_order_seq_cst);
	if (sleepers_.load(std::memory_order_seq_cst) > 0) {
		std::lock_guard<std::mutex> lk(mtx_);
		cv_.notify_all();
	}
}

void copy_pool::wait_piece(int p)
{
	if (p >= pieces_)
		p = pieces_ - 1;
	for (int q = 0; q <= p; q++)
		while (piece_done_[q].load(std::memory_order_acquire) < items_per_piece_) {
			if (!run_item())         /* help, one item at a time; nothing left to claim: wait for the claimants */
				CPU_RELAX();
		}
	if (p == pieces_ - 1 && open_.load(std::memory_order_relaxed)) {
		open_.store(false, std::memory_order_seq_cst);
		while (active_.load(std::memory_order_seq_cst) != 0)
			CPU_RELAX();
	}
}

void copy_pool::copy(void *dst, const void *src, size_t bytes)
{
	if (bytes == 0)
		return;
	if (bytes < 4 * MIN_ITEM) {             /* not worth waking anybody */
		memcpy(dst, src, bytes);
		return;
	}
	start(dst, src, bytes, 1);
	wait_piece(0);
}

} /* namespace fosphor_b200 */

/* C hook for the CPU test tier (tests/test_copy_pool.py): `rounds` jobs of `bytes` bytes in
 * `pieces` pieces on a pool of `threads` workers, every piece checked the moment wait_piece()
 * reports it, idle gaps long enough for the workers to fall asleep in between.  0 = all copies
 * exact. */
extern "C" int fosphor_host_copy_selftest(int threads, unsigned long long bytes, int pieces, int rounds)
{
	using fosphor_b200::copy_pool;
	if (bytes == 0 || rounds < 1)
		return -1;
	copy_pool pool(threads);
	std::vector<unsigned char> src(bytes), dst(bytes);
	unsigned long long state = 0x9e3779b97f4a7c15ull;
	for (int r = 0; r < rounds; r++) {
		for (size_t i = 0; i < bytes; i += 61) {
			state = state * 6364136223846793005ull + 1442695040888963407ull;
			src[i] = (unsigned char)(state >> 56);
		}
		memset(dst.data(), 0, bytes);
		pool.start(dst.data(), src.data(), bytes, pieces, (r & 1) != 0);
		for (int p = 0; p < pieces; p++) {
			pool.wait_piece(p);
			const size_t off = pool.piece_offset(p);
			if (off < bytes && memcmp(dst.data() + off, src.data() + off, pool.piece_size(p)) != 0)
				return 1 + r;
		}
		if (memcmp(dst.data(), src.data(), bytes) != 0)
			return 1 + r;
		if (r % 7 == 3)
			std::this_thread::sleep_for(std::chrono::milliseconds(6));   /* past the polling window: workers sleep */
		pool.copy(dst.data(), src.data(), bytes / 3 + 1);
	}
	return 0;
}
