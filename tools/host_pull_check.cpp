// MAS_OPT_HOST_PULL check without Python: an N x N spring cloth built here, set up once, then the host-pointer
// mas_apply on PAGE-LOCKED r / z with the copy-engine staging (option 0) and the kernel pull (option 1): the two z must be
// bit-identical; prints the wall-clock time per apply of both.
//   g++ -O2 -std=c++17 -I include -I /usr/local/cuda/include tools/host_pull_check.cpp -L <pkg> -lmas_b200 \
//       -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,<pkg> -o tools/host_pull_check
//   tools/host_pull_check [N=1024] [reps=20]
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime_api.h>

#include "mas_b200.h"

static double now()
{
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv)
{
	const int N = argc > 1 ? atoi(argv[1]) : 1024, reps = argc > 2 ? atoi(argv[2]) : 20;
	const int nv = N * N;
	const double t0 = now();
	std::vector<float> pos(4 * (size_t)nv, 0.f), diag(9 * (size_t)nv, 0.f);
	std::vector<int> starts(nv + 1, 0), idx;
	std::vector<float> off;
	idx.reserve(8 * (size_t)nv);
	off.reserve(72 * (size_t)nv);
	const float k = 1000.f;
	for (int v = 0; v < nv; ++v)
	{
		const int i = v % N, j = v / N;
		pos[4 * (size_t)v] = 0.01f * i;
		pos[4 * (size_t)v + 1] = 0.01f * j;
		float D[9] = { 1.f, 0, 0, 0, 1.f, 0, 0, 0, 1.f };
		for (int dj = -1; dj <= 1; ++dj)
			for (int di = -1; di <= 1; ++di)
			{
				if (!di && !dj) continue;
				const int ii = i + di, jj = j + dj;
				if (ii < 0 || jj < 0 || ii >= N || jj >= N) continue;
				const float len = std::sqrt((float)(di * di + dj * dj));
				const float d[3] = { di / len, dj / len, 0.f };
				float K[9];
				for (int a = 0; a < 3; ++a)
					for (int b = 0; b < 3; ++b) K[3 * b + a] = k * d[a] * d[b] + (a == b ? 0.1f * k : 0.f);
				idx.push_back(jj * N + ii);
				for (int e = 0; e < 9; ++e) { off.push_back(-K[e]); D[e] += K[e]; }
			}
		std::memcpy(&diag[9 * (size_t)v], D, sizeof(D));
		starts[v + 1] = (int)idx.size();
	}
	mas_handle_t h = nullptr;
	if (mas_create(&h, 0) != MAS_OK) { printf("mas_create failed\n"); return 2; }
	if (mas_allocate(h, nv, 0, 0, pos.data(), nullptr, nullptr, starts.data(), idx.data(), MAS_MEM_HOST) != MAS_OK ||
		mas_prepare(h, diag.data(), off.data(), starts.data(), nullptr, nullptr, nullptr, 0, 0, 0, MAS_MEM_HOST) != MAS_OK)
	{
		printf("setup failed: %s\n", mas_last_error(h));
		return 3;
	}
	float *r = nullptr, *z0 = nullptr, *z1 = nullptr;
	const size_t bytes = 16 * (size_t)nv;
	if (cudaHostAlloc((void**)&r, bytes, cudaHostAllocDefault) != cudaSuccess || cudaHostAlloc((void**)&z0, bytes, cudaHostAllocDefault) != cudaSuccess ||
		cudaHostAlloc((void**)&z1, bytes, cudaHostAllocDefault) != cudaSuccess) { printf("cudaHostAlloc failed\n"); return 4; }
	unsigned s = 12345u;
	for (size_t q = 0; q < 4 * (size_t)nv; ++q)
	{
		s = s * 1664525u + 1013904223u;
		r[q] = (q & 3) == 3 ? 0.f : (float)(s >> 8) / 8388608.f - 1.f;
	}
	std::memset(z0, 0x7f, bytes);
	std::memset(z1, 0x3f, bytes);
	const double t1 = now();
	double ms[2] = { 0, 0 };
	float* zs[2] = { z0, z1 };
	for (int mode = 0; mode < 2; ++mode)
	{
		if (mas_set_option(h, MAS_OPT_HOST_PULL, mode) != MAS_OK) { printf("set_option failed\n"); return 5; }
		for (int w = 0; w < 3; ++w)
			if (mas_apply(h, zs[mode], r, MAS_MEM_HOST) != MAS_OK) { printf("apply failed (mode %d): %s\n", mode, mas_last_error(h)); return 6; }
		const double a = now();
		for (int q = 0; q < reps; ++q) mas_apply(h, zs[mode], r, MAS_MEM_HOST);
		ms[mode] = (now() - a) * 1e3 / reps;
	}
	const int same = std::memcmp(z0, z1, bytes) == 0;
	double nrm = 0;
	for (size_t q = 0; q < 4 * (size_t)nv; ++q) nrm += (double)z1[q] * z1[q];
	printf("{\"nv\": %d, \"identical\": %d, \"z_norm\": %.6e, \"copy_engine_ms\": %.4f, \"host_pull_ms\": %.4f, \"build_s\": %.2f, \"total_s\": %.2f}\n", nv, same,
		std::sqrt(nrm), ms[0], ms[1], t1 - t0, now() - t0);
	mas_destroy(h);
	return same ? 0 : 1;
}
