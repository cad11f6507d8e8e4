/*
 * copy_pool.h - a few persistent threads that move sample / result buffers
 * between the caller's ordinary (pageable) memory and the engine's
 * page-locked staging memory.
 *
 * Why it exists: the reference sink hands fosphor_cl_process() pointers into a
 * pageable 16 MiB ring (lib/fifo.cc:17-21, lib/base_sink_c_impl.cc:58,169-170)
 * and may recycle the region as soon as the call returns (:174), so the call
 * has to take its own copy.  One thread copies ~12 GB/s on the GPU box, PCIe
 * moves 53 GB/s: the copy, not the bus, would set the pace (round 1: 1.3
 * Gsamples/s).  Eight threads copy ~80 GB/s, and because the job is cut into
 * pieces that complete IN ORDER the caller can start the DMA of piece 0 while
 * the threads are on piece 1.
 *
 * Workers spin briefly for the next job (a streaming caller arrives every
 * ~150 us) and then sleep on a condition variable, so an idle engine costs no
 * CPU and an oversubscribed host degrades instead of collapsing.
 */
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <mutex>
#include <thread>
#include <vector>

namespace fosphor_b200 {

class copy_pool {
public:
	static constexpr int MAX_PIECES = 64;

	/* n_threads <= 0: automatic (see copy_pool.cc: default_threads) */
	explicit copy_pool(int n_threads);
	~copy_pool();
	copy_pool(const copy_pool &) = delete;
	copy_pool &operator=(const copy_pool &) = delete;

	int threads() const { return (int)workers_.size(); }

	/* Start copying bytes from src to dst, cut into `pieces` equal parts (the last one takes the
	 * remainder; 1 <= pieces <= MAX_PIECES) that complete in order.  Returns at once; the calling
	 * thread helps from wait_piece().  One job at a time.  non_temporal: write the destination
	 * around the caches (staging buffers the DMA engine reads next). */
	void start(void *dst, const void *src, size_t bytes, int pieces, bool non_temporal = false);
	/* Block until piece p (and all before it) has been copied. */
	void wait_piece(int p);
	/* start + wait for everything */
	void copy(void *dst, const void *src, size_t bytes);

	size_t piece_offset(int p) const { return (size_t)p * piece_bytes_; }
	size_t piece_size(int p) const
	{
		return p + 1 < pieces_ ? piece_bytes_ : bytes_ - (size_t)(pieces_ - 1) * piece_bytes_;
	}

	static int default_threads();

private:
	void worker();
	bool run_item();           /* claim and copy one work item; false when the job has none left */

	std::vector<std::thread> workers_;
	std::mutex mtx_;
	std::condition_variable cv_;
	std::atomic<unsigned> generation_{0};
	std::atomic<int> sleepers_{0};
	std::atomic<bool> stop_{false};
	std::atomic<bool> open_{false};      /* a job is published and not yet closed */
	std::atomic<int> active_{0};         /* workers that may be inside run_item() */

	/* the current job */
	char *dst_ = nullptr;
	const char *src_ = nullptr;
	size_t bytes_ = 0, piece_bytes_ = 0, item_bytes_ = 0;
	int pieces_ = 0, items_per_piece_ = 0, n_items_ = 0;
	bool nt_ = false;                    /* non-temporal stores for this job */
	std::atomic<int> next_item_{0};
	std::atomic<int> piece_done_[MAX_PIECES];
};

} /* namespace fosphor_b200 */
