// A small persistent pool of host threads for the per-frame host work of many-mesh scenes.
//
// The reference's Renderer::render() walks the scene graph and forms one modelview / normal matrix per renderable
// (src/Renderer.cpp:311-338) on one thread; with 10 000 meshes (BASELINE.json configs[3]) that serial walk costs
// more host time than the whole frame takes on the device. Every entry is independent, so the walk is dealt to a
// few threads that live as long as the library: starting threads per frame was measured slower than one thread.
// The arithmetic of an entry is the same code whichever thread runs it: results do not depend on the width.
//
// Only library code runs on the workers (never an application's virtual overrides). The pool is created on first
// use by a frame with enough entries, re-created in a forked child, and joined when the library is unloaded.
// MINIRENDER_B200_HOST_THREADS=1 turns it off, =N fixes the width (default: min(8, hardware threads / 2)).
// Jobs must not throw and must not start jobs themselves (one job at a time: callers take turns).
#ifndef MINIRENDER_B200_HOST_POOL_H
#define MINIRENDER_B200_HOST_POOL_H

#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <mutex>
#include <new>
#include <thread>
#include <vector>
#include <pthread.h>

namespace minirender {
namespace hostpool {

class Pool
{
public:
	typedef void (*Job)(void* arg, int chunk);

	// The process-wide pool, or 0 when the width is 1.
	static Pool* get()
	{
		static Holder holder;
		Pool* p = holder.pool.load(std::memory_order_acquire);
		if (!p && !holder.off)
		{
			std::lock_guard<std::mutex> lock(holder.m);
			p = holder.pool.load(std::memory_order_acquire);
			if (!p)
			{
				// half of what the machine reports, at most 8: hardware threads are often two per core, and on a
				// host whose processors are shared with others a worker that loses its processor holds a frame up
				int width = (int)std::thread::hardware_concurrency() / 2;
				if (width > 8)
					width = 8;
				const char* e = getenv("MINIRENDER_B200_HOST_THREADS");
				if (e && atoi(e) > 0)
					width = atoi(e);
				if (width > 64)
					width = 64;
				if (width <= 1)
				{
					holder.off = true;
					return 0;
				}
				if (!holder.forkHooked)
				{
					pthread_atfork(0, 0, &Pool::afterForkInChild);
					holder.forkHooked = true;
				}
				slot() = &holder;
				p = new Pool(width);
				holder.pool.store(p, std::memory_order_release);
			}
		}
		return p;
	}

	int width() const { return (int)_workers.size() + 1; }

	// Runs job(arg, c) for every c in [0, chunks) on the workers and the calling thread; returns when all are done.
	// One caller at a time (frames of different Renderer objects take turns).
	void run(Job job, void* arg, int chunks)
	{
		if (chunks <= 0)
			return;
		std::lock_guard<std::mutex> turn(_callers);
		unsigned mine;
		{
			std::lock_guard<std::mutex> lock(_m);
			_job = job;
			_arg = arg;
			_chunks = chunks;
			_left.store(chunks, std::memory_order_relaxed);
			mine = _generation.load(std::memory_order_relaxed) + 1u;
			_next.store((unsigned long long)mine << 32, std::memory_order_relaxed);
			_generation.store(mine, std::memory_order_release);
		}
		_wake.notify_all();
		work(job, arg, chunks, mine);
		// the last chunks may still be running on workers
		for (int spin = 0; _left.load(std::memory_order_acquire) > 0; spin++)
		{
			relax();
			if (spin > 4000)
				std::this_thread::yield();
		}
	}

private:
	struct Holder
	{
		std::atomic<Pool*> pool;
		std::mutex m;
		bool off, forkHooked;
		Holder() : pool(0), off(false), forkHooked(false) {}
		~Holder()
		{
			Pool* p = pool.exchange(0);
			delete p;
		}
	};
	static Holder*& slot()
	{
		static Holder* h = 0;
		return h;
	}
	// A forked child has this object but none of its threads: forget it (a new pool is made on demand).
	static void afterForkInChild()
	{
		Holder* h = slot();
		if (h)
		{
			h->pool.store(0, std::memory_order_release); // the parent's object is leaked in the child on purpose
			new (&h->m) std::mutex();
		}
	}

	static inline void relax()
	{
#if defined(__x86_64__) || defined(__i386__)
		__builtin_ia32_pause();
#endif
	}

	explicit Pool(int width) : _job(0), _arg(0), _chunks(0), _next(0), _left(0), _generation(0), _quit(false)
	{
		for (int i = 1; i < width; i++)
			_workers.push_back(std::thread(&Pool::loop, this));
	}
	~Pool()
	{
		{
			std::lock_guard<std::mutex> lock(_m);
			_quit = true;
			_generation.fetch_add(1u, std::memory_order_release);
		}
		_wake.notify_all();
		for (size_t i = 0; i < _workers.size(); i++)
			_workers[i].join();
	}

	// Tickets carry the job's generation: a worker that wakes up late, with the fields of a job that is already
	// over, finds another generation in the counter and takes nothing.
	void work(Job job, void* arg, int chunks, unsigned generation)
	{
		unsigned long long v = _next.load(std::memory_order_relaxed);
		for (;;)
		{
			if ((unsigned)(v >> 32) != generation || (int)(v & 0xffffffffull) >= chunks)
				break;
			if (!_next.compare_exchange_weak(v, v + 1ull, std::memory_order_acq_rel, std::memory_order_relaxed))
				continue;
			job(arg, (int)(v & 0xffffffffull));
			_left.fetch_sub(1, std::memory_order_release);
			v = _next.load(std::memory_order_relaxed);
		}
	}

	void loop()
	{
		unsigned seen = 0;
		for (;;)
		{
			// a frame runs two or three jobs back to back: poll for a moment before sleeping
			for (int spin = 0; spin < 1000 && _generation.load(std::memory_order_acquire) == seen; spin++)
				relax();
			Job job;
			void* arg;
			int chunks;
			{
				std::unique_lock<std::mutex> lock(_m);
				while (_generation.load(std::memory_order_acquire) == seen)
					_wake.wait(lock);
				seen = _generation.load(std::memory_order_acquire);
				if (_quit)
					return;
				job = _job;
				arg = _arg;
				chunks = _chunks;
			}
			work(job, arg, chunks, seen);
		}
	}

	std::vector<std::thread> _workers;
	std::mutex _m, _callers;
	std::condition_variable _wake;
	Job _job;
	void* _arg;
	int _chunks;
	std::atomic<unsigned long long> _next; // job generation << 32 | next chunk
	std::atomic<int> _left;
	std::atomic<unsigned> _generation;
	bool _quit;
};

// Runs f(c) for c in [0, chunks): on the pool when there is one, in order on the caller otherwise.
template <class F>
static inline void parallelFor(Pool* pool, int chunks, F& f)
{
	struct Thunk
	{
		static void call(void* arg, int c) { (*(F*)arg)(c); }
	};
	if (pool && chunks > 1)
		pool->run(&Thunk::call, (void*)&f, chunks);
	else
		for (int c = 0; c < chunks; c++)
			f(c);
}

}
}
#endif
