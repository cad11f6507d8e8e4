/*
 * pinned_fifo.cc - see pinned_fifo.h.  One slot is always kept empty so that
 * wp == rp means "empty" (as in the reference ring, lib/fifo.cc:28-38).
 */
#include "pinned_fifo.h"

#include <cstdlib>
#include <new>

#include <cuda_runtime.h>

namespace fosphor_b200 {

pinned_fifo::pinned_fifo(int length)
	: buf_(nullptr), len_(length), mask_(length - 1), rp_(0), wp_(0), pinned_(false)
{
	if (length < 2 || (length & (length - 1)))
		return;
	void *p = nullptr;
	if (cudaHostAlloc(&p, sizeof(sample) * (size_t)length, cudaHostAllocPortable) == cudaSuccess) {
		pinned_ = true;
	} else {
		cudaGetLastError();
		p = std::malloc(sizeof(sample) * (size_t)length);
	}
	buf_ = static_cast<sample *>(p);
}

pinned_fifo::~pinned_fifo()
{
	if (!buf_)
		return;
	if (pinned_)
		cudaFreeHost(buf_);
	else
		std::free(buf_);
}

int pinned_fifo::used()
{
	return (wp_ - rp_) & mask_;
}

int pinned_fifo::free()
{
	return mask_ - used();
}

int pinned_fifo::write_max_size()
{
	return len_ - wp_;
}

int pinned_fifo::read_max_size()
{
	return len_ - rp_;
}

pinned_fifo::sample *pinned_fifo::write_prepare(int size, bool wait)
{
	std::unique_lock<std::mutex> lk(mtx_);
	if (!wait && free() < size)
		return nullptr;
	not_full_.wait(lk, [&] { return free() >= size; });
	return buf_ + wp_;
}

void pinned_fifo::write_commit(int size)
{
	{
		std::lock_guard<std::mutex> lk(mtx_);
		wp_ = (wp_ + size) & mask_;
	}
	not_empty_.notify_one();
}

pinned_fifo::sample *pinned_fifo::read_peek(int size, bool wait)
{
	std::unique_lock<std::mutex> lk(mtx_);
	if (!wait && used() < size)
		return nullptr;
	not_empty_.wait(lk, [&] { return used() >= size; });
	return buf_ + rp_;
}

void pinned_fifo::read_discard(int size)
{
	{
		std::lock_guard<std::mutex> lk(mtx_);
		rp_ = (rp_ + size) & mask_;
	}
	not_full_.notify_one();
}

} /* namespace fosphor_b200 */

/* ---- plain C handle for non-C++ callers and the tests -------------------- */
extern "C" {

void *fosphor_fifo_create(int length)
{
	auto *f = new (std::nothrow) fosphor_b200::pinned_fifo(length);
	if (f && !f->ok()) {
		delete f;
		f = nullptr;
	}
	return f;
}

void fosphor_fifo_destroy(void *h) { delete static_cast<fosphor_b200::pinned_fifo *>(h); }
int fosphor_fifo_is_pinned(void *h) { return static_cast<fosphor_b200::pinned_fifo *>(h)->pinned(); }
int fosphor_fifo_free(void *h) { return static_cast<fosphor_b200::pinned_fifo *>(h)->free(); }
int fosphor_fifo_used(void *h) { return static_cast<fosphor_b200::pinned_fifo *>(h)->used(); }
int fosphor_fifo_write_max_size(void *h) { return static_cast<fosphor_b200::pinned_fifo *>(h)->write_max_size(); }
int fosphor_fifo_read_max_size(void *h) { return static_cast<fosphor_b200::pinned_fifo *>(h)->read_max_size(); }

void *fosphor_fifo_write_prepare(void *h, int size, int wait)
{
	return static_cast<fosphor_b200::pinned_fifo *>(h)->write_prepare(size, wait != 0);
}

void fosphor_fifo_write_commit(void *h, int size) { static_cast<fosphor_b200::pinned_fifo *>(h)->write_commit(size); }

void *fosphor_fifo_read_peek(void *h, int size, int wait)
{
	return static_cast<fosphor_b200::pinned_fifo *>(h)->read_peek(size, wait != 0);
}

void fosphor_fifo_read_discard(void *h, int size) { static_cast<fosphor_b200::pinned_fifo *>(h)->read_discard(size); }

} /* extern "C" */
