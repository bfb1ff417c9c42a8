// Stress of minirender_b200/host/HostPool.h on its own (no renderer): many short jobs back to back, jobs of one chunk
// and of thousands, several application threads calling at once, pauses long enough for the workers to fall asleep.
// Every chunk of every job must run exactly once. Prints "OK <jobs>"; meant to be built with and without
// -fsanitize=thread.
#include "../../minirender_b200/host/HostPool.h"
#include <chrono>
#include <cstdio>
#include <cstring>

using namespace minirender::hostpool;

struct Job
{
	std::vector<int>* hits;
	std::atomic<long long>* sum;
	void operator()(int c)
	{
		(*hits)[(size_t)c]++;
		sum->fetch_add(c + 1, std::memory_order_relaxed);
	}
};

static int caller(Pool* pool, unsigned seed, int jobs, std::atomic<int>* bad)
{
	for (int j = 0; j < jobs; j++)
	{
		seed = seed * 1664525u + 1013904223u;
		const int chunks = (j % 7 == 0) ? 1 : (j % 11 == 0) ? 3000 : 1 + (int)((seed >> 10) % 97);
		std::vector<int> hits((size_t)chunks, 0);
		std::atomic<long long> sum(0);
		Job job = { &hits, &sum };
		parallelFor(pool, chunks, job);
		for (int c = 0; c < chunks; c++)
			if (hits[(size_t)c] != 1)
				bad->fetch_add(1);
		if (sum.load() != (long long)chunks * (chunks + 1) / 2)
			bad->fetch_add(1);
		if (j % 50 == 49)
			std::this_thread::sleep_for(std::chrono::milliseconds(2)); // the workers go to sleep
	}
	return 0;
}

int main(int argc, char** argv)
{
	const int jobs = argc > 1 ? atoi(argv[1]) : 2000;
	Pool* pool = Pool::get();
	if (!pool)
	{
		printf("OK 0 (no pool: one hardware thread or MINIRENDER_B200_HOST_THREADS=1)\n");
		return 0;
	}
	std::atomic<int> bad(0);
	caller(pool, 1u, jobs, &bad);
	std::vector<std::thread> callers;
	for (int t = 0; t < 3; t++)
		callers.push_back(std::thread(caller, pool, 100u + t, jobs / 2, &bad));
	for (size_t t = 0; t < callers.size(); t++)
		callers[t].join();
	if (bad.load())
	{
		printf("FAIL %d\n", bad.load());
		return 1;
	}
	printf("OK %d\n", jobs + 3 * (jobs / 2));
	return 0;
}
