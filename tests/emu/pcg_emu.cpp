// TEST INFRASTRUCTURE ONLY.  The kernels of the device PCG harness (csrc/mas_pcg.cu: sliced-ELL conversion, spmv_dot, axpy_rr,
// dot, update_p with its device-side stopping test) run on the CPU through tests/emu/cuda_emu.h, in the launch order of
// pcg_solve() WITHOUT preconditioner (z = r: plain CG) for a fixed number of iterations.
//   pcg_emu < in.bin > out.bin
//   in : int32 nv, nnz, iterations; float32 tol; int32 ranges[nv + 1], idx[nnz]; float32 diag[nv][9], offdiag[nnz][9], b[nv][4]
//   out: float32 x[nv][4], r[nv][4], Ap[nv][4] (of the last iteration's SpMV); float64 rr, rr0; int32 iters, done
#include "cuda_emu.h"

#include <cstdio>
#include <cstdlib>
#include <numeric>

#define MAS_CPU_EMULATION 1
#include "../../preconditioner-for-cloth-and-deformable-body-simulation_b200/csrc/mas_pcg.cu"

template <typename T>
static std::vector<T> rd(size_t n)
{
	std::vector<T> v(n ? n : 1);
	if (n && fread(v.data(), sizeof(T), n, stdin) != n) { fprintf(stderr, "short read\n"); exit(2); }
	return v;
}

int main()
{
	using namespace mas;
	const auto hdr = rd<int>(3);
	const int nv = hdr[0], nnz = hdr[1], iterations = hdr[2];
	const float relTol = rd<float>(1)[0];
	const auto ranges = rd<int>((size_t)nv + 1);
	const auto idx = rd<int>((size_t)nnz);
	const auto diag = rd<float>((size_t)nv * 9);
	const auto off = rd<float>((size_t)nnz * 9);
	const auto b = rd<float4>((size_t)nv);

	std::vector<float4> x((size_t)nv), r((size_t)nv), p((size_t)nv), Ap((size_t)nv);
	std::vector<double> partials((size_t)3 * kMaxPartials, 0.0);
	double* pA = partials.data();
	double* pRR = pA + kMaxPartials;
	double* pRZ = pRR + kMaxPartials;
	PcgState state = {};
	float4* z = r.data();                                   // usePrecond = 0
	// fewer CTAs than rows / 256 (as the occupancy-sized launches of pcg_solve have on large meshes): the kernels' strided
	// trips, the reloads inside them and the ragged last trip are all exercised
	int grid = cdiv(cdiv(nv, kPcgThreads), 3);
	if (grid < 1) grid = 1;
	int gridSpmv = cdiv(cdiv(cdiv(nv, 32), kPcgWarps), 3);
	if (gridSpmv < 1) gridSpmv = 1;
	const int nPart = grid > gridSpmv ? grid : gridSpmv;
	const double tol2 = (double)relTol * (double)relTol;

	const int nSlices = cdiv(nv, 32);
	std::vector<int> sliceSlots((size_t)nSlices), sliceStart((size_t)nSlices);
	emu::launch(cdiv((long long)nSlices * 32, 256), 256, [&] { ell_width_kernel(ranges.data(), nv, sliceSlots.data()); });
	std::exclusive_scan(sliceSlots.begin(), sliceSlots.end(), sliceStart.begin(), 0);      // launch_exclusive_scan
	const int totalSlots = nSlices ? sliceStart.back() + sliceSlots.back() : 0;
	std::vector<int> ellIdx((size_t)(totalSlots > 0 ? totalSlots : 1));
	std::vector<float> ellVal((size_t)(totalSlots > 0 ? totalSlots : 1) * 9);
	emu::launch(cdiv((long long)nSlices * 32, 256), 256, [&] {
		ell_fill_kernel(diag.data(), off.data(), ranges.data(), idx.data(), nv, sliceStart.data(), sliceSlots.data(), ellIdx.data(), ellVal.data());
	});

	emu::launch(cdiv(nv, 256), 256, [&] { copy_b_kernel(b.data(), r.data(), x.data(), nv); });
	emu::launch(grid, kPcgThreads, [&] { dot_kernel(r.data(), r.data(), nv, pRR, &state); });
	emu::launch(grid, kPcgThreads, [&] { dot_kernel(r.data(), z, nv, pRZ, &state); });
	emu::launch(grid, kPcgThreads, [&] { update_p_kernel(x.data(), p.data(), z, nv, pRZ, pRR, nPart, tol2, iterations, 0, &state); });
	for (int it = 0; it < iterations; ++it)
	{
		emu::launch(gridSpmv, kPcgThreads, [&] {
			spmv_dot_kernel(sliceStart.data(), sliceSlots.data(), ellIdx.data(), ellVal.data(), p.data(), Ap.data(), nv, pA, &state);
		});
		emu::launch(grid, kPcgThreads, [&] { axpy_rr_kernel(r.data(), Ap.data(), nv, pA, nullptr, nPart, pRR, &state); });
		emu::launch(grid, kPcgThreads, [&] { dot_kernel(r.data(), z, nv, pRZ, &state); });
		emu::launch(grid, kPcgThreads, [&] { update_p_kernel(x.data(), p.data(), z, nv, pRZ, pRR, nPart, tol2, iterations, 1, &state); });
	}
	fwrite(x.data(), 16, (size_t)nv, stdout);
	fwrite(r.data(), 16, (size_t)nv, stdout);
	fwrite(Ap.data(), 16, (size_t)nv, stdout);
	fwrite(&state.rr, 8, 1, stdout);
	fwrite(&state.rr0, 8, 1, stdout);
	fwrite(&state.iters, 4, 1, stdout);
	fwrite(&state.done, 4, 1, stdout);
	return 0;
}
