// TEST INFRASTRUCTURE ONLY.  Runs the FP32 CUDA-core inversion of csrc/mas_invert.cuh (MAS_OPT_INVERT_VARIANT = 1) — the very
// text the CUDA kernel compiles — on the CPU (tests/emu/cuda_emu.h), on matrices read from stdin, and writes the unpacked
// inverses to stdout.  (The default tensor-core kernel, mas_invert_tc.cuh, needs tcgen05 hardware; its algorithm is replayed
// in numpy by tools/sweep_inversion_study.py.)
//   invert_emu < in.bin > out.bin      in: int32 count, then count x 96 x 96 float32 (row-major, symmetric)
//                                      out: count x 96 x 96 float32
#include "cuda_emu.h"

#include <cstddef>
#include <cstdio>
#include <cstdlib>

#define MAS_CPU_EMULATION 1
#include "../../preconditioner-for-cloth-and-deformable-body-simulation_b200/csrc/mas_internal.h"

namespace mas {
namespace {
struct PhaseClock
{
	void start() {}
	void mark(int) {}
};
#include "../../preconditioner-for-cloth-and-deformable-body-simulation_b200/csrc/mas_invert.cuh"

void invert_one(const float* dense, float* out)
{
	std::vector<unsigned char> smem(sizeof(InvSmem) + 64);
	unsigned char* base = smem.data() + ((16 - (reinterpret_cast<uintptr_t>(smem.data()) & 15)) & 15);
	InvSmem& s = *reinterpret_cast<InvSmem*>(base);
	for (int r = 0; r < kDof; ++r)
		for (int c = 0; c < kDof; ++c) s.A[tile_at(r, c)] = dense[r * kDof + c];
	// the table ensure_pos_table() builds on the host (csrc/mas_assemble.cu)
	std::vector<unsigned short> posTab((size_t)kOutPerThread * kInvThreads);
	for (int t = 0; t < kInvThreads; ++t)
	{
		const int tr = t & 15, tc = t >> 4;
		int e = 0;
		for (int i = 0; i < 6; ++i)
			for (int j = 0; j < i; ++j, ++e) posTab[(size_t)e * kInvThreads + t] = (unsigned short)packed_pos(tr + 16 * i, tc + 16 * j);
		for (int i = 0; i < 6; ++i)
			posTab[(size_t)(15 + i) * kInvThreads + t] = (unsigned short)(tr >= tc ? packed_pos(tr + 16 * i, tc + 16 * i) : 0);
	}
	std::vector<float> packed(kTri);
	emu::run(kInvThreads, [&] {
		PhaseClock pc;
		const float* p = invert_tile(s, posTab.data(), pc);
		store_packed(p, packed.data());
	});
	for (int r = 0; r < kDof; ++r)
		for (int c = 0; c < kDof; ++c) out[r * kDof + c] = packed[packed_pos(r, c)];
}
}  // namespace
}  // namespace mas

int main(int argc, char** argv)
{
	(void)argc; (void)argv;
	int count = 0;
	if (fread(&count, 4, 1, stdin) != 1 || count < 0 || count > 4096) return 2;
	std::vector<float> in((size_t)count * 96 * 96), out((size_t)count * 96 * 96);
	if (fread(in.data(), 4, in.size(), stdin) != in.size()) return 2;
	for (int b = 0; b < count; ++b)
	{
		const float* src = in.data() + (size_t)b * 9216;
		float* dst = out.data() + (size_t)b * 9216;
		mas::invert_one(src, dst);
	}
	fwrite(out.data(), 4, out.size(), stdout);
	return 0;
}
