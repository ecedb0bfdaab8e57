// TEST INFRASTRUCTURE ONLY — never linked into the product library.
//
// C-ABI harness around the UNMODIFIED older variant of the reference,
// SE::SeSchwarzPreconditionerPreviousVersion (SeSchwarzPreconditionerPreviousVersion.h:39-128, header-only), compiled by
// oracle/build_ref.sh into oracle/_ref/libmas_prev.so.  Same public surface as the current class (h:46-58, 87, 118); used as
// a SECOND cross-check of z (identical ordering, z differs from the current version by rounding only).
#include <cstring>
#include <vector>

#define private public
#include "SeSchwarzPreconditionerPreviousVersion.h"
#undef private

using namespace SE;

extern int CPU_THREAD_NUM;  // SeOmp.cpp

namespace {
struct PrevHandle
{
	SeSchwarzPreconditionerPreviousVersion pre;
	SeCsr<int>* csr = nullptr;
	std::vector<SeVec3fSimd> positions;
	std::vector<Int4> edges, faces;
	std::vector<unsigned int> efCounts, eeCounts, vfCounts;
	int nv = 0, ne = 0, nf = 0;
};
}

extern "C" {

void prev_set_threads(int n) { CPU_THREAD_NUM = n > 0 ? n : 1; }

void* prev_create() { return new PrevHandle(); }

void prev_destroy(void* h)
{
	auto* r = (PrevHandle*)h;
	delete r->csr;
	delete r;
}

void prev_allocate(void* h, int nv, int ne, int nf, const float* positions, const int* edges, const int* faces,
	const int* nbrStarts, const int* nbrIdx)
{
	auto* r = (PrevHandle*)h;
	r->nv = nv; r->ne = ne; r->nf = nf;
	r->positions.resize(nv);
	std::memcpy((void*)r->positions.data(), positions, sizeof(float) * 4 * (size_t)nv);
	r->edges.resize(ne > 0 ? ne : 1);
	r->faces.resize(nf > 0 ? nf : 1);
	if (ne > 0) std::memcpy((void*)r->edges.data(), edges, sizeof(int) * 4 * (size_t)ne);
	if (nf > 0) std::memcpy((void*)r->faces.data(), faces, sizeof(int) * 4 * (size_t)nf);
	std::vector<int> starts(nbrStarts, nbrStarts + nv + 1);
	std::vector<int> idxs(nbrIdx, nbrIdx + starts.back());
	std::vector<int> values;
	delete r->csr;
	r->csr = new SeCsr<int>(starts, idxs, values);
	r->pre.m_positions = r->positions.data();
	r->pre.m_edges = r->edges.data();
	r->pre.m_faces = r->faces.data();
	r->pre.m_neighbours = r->csr;
	r->pre.AllocatePrecoditioner(nv, ne, nf);
}

void prev_prepare(void* h, const float* diag, const float* offdiag, const int* ranges,
	const void* ef, const void* ee, const void* vf, unsigned efTotal, unsigned eeTotal, unsigned vfTotal)
{
	auto* r = (PrevHandle*)h;
	r->efCounts.assign((size_t)r->ne + 1, 0u); r->efCounts[r->ne] = efTotal;
	r->eeCounts.assign((size_t)r->ne + 1, 0u); r->eeCounts[r->ne] = eeTotal;
	r->vfCounts.assign((size_t)r->nv + 1, 0u); r->vfCounts[r->nv] = vfTotal;
	r->pre.PreparePreconditioner((const SeMatrix3f*)diag, (const SeMatrix3f*)offdiag, ranges,
		(const EfSet*)ef, (const EeSet*)ee, (const VfSet*)vf,
		r->efCounts.data(), r->eeCounts.data(), r->vfCounts.data());
}

void prev_apply(void* h, float* z, const float* residual)
{
	auto* r = (PrevHandle*)h;
	r->pre.Preconditioning((SeVec3fSimd*)z, (const SeVec3fSimd*)residual, 3 * r->nv);
}

void prev_get_sorted_get_original(void* h, int* out)
{
	auto& p = ((PrevHandle*)h)->pre;
	std::memcpy(out, p.m_MapperSortedGetOriginal.data(), sizeof(int) * (size_t)p.m_numVerts);
}

int prev_total_clusters(void* h) { return ((PrevHandle*)h)->pre.m_totalNumberClusters; }

}  // extern "C"
