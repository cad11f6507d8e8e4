/*
 * pinned_fifo.h - page-locked sample FIFO for the sink side of the boundary.
 *
 * SURVEY.md section 8(f) #2.  Same role and method names as the reference's
 * gr::fosphor::fifo (lib/fifo.h:20-46: a single-producer / single-consumer ring
 * of complex samples between base_sink_c_impl::work() and the render thread,
 * lib/base_sink_c_impl.cc:432-462,146-175), with the two changes the CUDA
 * engine wants:
 *   - the storage is page-locked (cudaHostAlloc), so fosphor_cl_process() /
 *     fosphor_cu_process_host() DMA straight out of the ring instead of staging
 *     through a CPU memcpy (engine.cu: upload_staged);
 *   - no GNU Radio dependency (std::mutex / std::condition_variable).
 * When no CUDA device is usable the ring falls back to ordinary memory
 * (pinned() == false); the engine then stages as for any pageable buffer.
 */
#pragma once
#include <complex>
#include <condition_variable>
#include <cstddef>
#include <mutex>

namespace fosphor_b200 {

class pinned_fifo {
public:
	typedef std::complex<float> sample;

	explicit pinned_fifo(int length);        /* length: power of two, in samples */
	~pinned_fifo();
	pinned_fifo(const pinned_fifo &) = delete;
	pinned_fifo &operator=(const pinned_fifo &) = delete;

	bool ok() const { return buf_ != nullptr; }
	bool pinned() const { return pinned_; }
	int length() const { return len_; }

	int free();                              /* samples that can still be written */
	int used();                              /* samples waiting to be read        */

	int write_max_size();                    /* contiguous space up to the ring end */
	sample *write_prepare(int size, bool wait = true);
	void write_commit(int size);

	int read_max_size();                     /* contiguous data up to the ring end */
	sample *read_peek(int size, bool wait = true);
	void read_discard(int size);

private:
	sample *buf_;
	int len_, mask_;
	int rp_, wp_;
	bool pinned_;
	std::mutex mtx_;
	std::condition_variable not_empty_, not_full_;
};

} /* namespace fosphor_b200 */
