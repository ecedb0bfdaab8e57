/* TEST INFRASTRUCTURE ONLY — CPU restatement of the MAS preconditioner path.
 *
 * Plain-C restatement of what SE::SeSchwarzPreconditioner computes
 * (/root/reference/SeSchwarzPreconditioner.cpp, cited per function as cpp:LINE).
 * It is the checker for the CUDA library: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline/--impl reference legs may load it.  The product
 * (libmas_b200.so) never links or calls anything in oracle/.
 *
 * Parity pin: tests/test_oracle_vs_reference.py compares every integer
 * structure produced here bit-for-bit, and every FP result by tolerance,
 * against the reference's own code compiled into oracle/_ref/libmas_ref.so
 * (oracle/build_ref.sh), plus the committed fixtures under tests/golden/.
 * The reference ships no tests or golden vectors of its own.
 *
 * The file is compiled twice (oracle/Makefile): REAL=float gives maso_f_*
 * (same arithmetic type as the reference) and REAL=double gives maso_d_*,
 * the FP64 arbiter used to state tolerances.
 *
 * Deliberate differences from the reference, all documented in DESIGN.md:
 *  - equal Morton codes are ordered by ascending original index (std::sort at
 *    cpp:242 leaves ties unspecified);
 *  - cluster ids come from a correct exclusive scan (bug Q5 at cpp:989-994 is
 *    not reproduced; oracle/_ref/libmas_ref_q5fix.so is the comparison there);
 *  - buffers are sized from the actual cluster counts (Q6);
 *  - stencils are compacted in input order (the atomic counter at cpp:407
 *    gives that order with one thread);
 *  - the "remaining neighbour" lists (cpp:74-75, 486-491, 788-793) are not
 *    materialised: an edge consumed at one level joins both ends into one
 *    cluster, so at later levels it can only set a node's own bit, which
 *    NextLevelCluster sets anyway (cpp:898).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef REAL
#define REAL float
#endif
#ifndef PFX
#define PFX maso_f_
#endif
#define CAT2(a, b) a##b
#define CAT(a, b) CAT2(a, b)
#define FN(name) CAT(PFX, name)

#define BANK 32
#define DOF 96
#define TRI 4656 /* 96*97/2 */

typedef struct
{
	int n, nFirst;
	int index[5];
	float weight[5];
	float stiff;
	float dir[4];
} Stencil; /* SeCollisionElements.h:60-69 field order */

typedef struct
{
	int nv, ne, nf, nnz;
	int numLevel;
	int prolongAllLevels; /* 0 = reproduce Q4 (cpp:1710) */
	float lo[4], hi[4];
	uint64_t* code;
	int *s2o, *o2s;
	int *nbrStart, *nbrIdx, *nbrSrc; /* sorted-space adjacency; nbrSrc = index of the 3x3 block in csrOffDiagonals */
	int *edges, *faces;
	/* prepare */
	int nStencil;
	Stencil* st;
	int (*stIdx)[5];
	uint32_t* fineMask;
	int* cst[8]; /* CoarseSpaceTables[level][v] (h:88) */
	int* goingNext;
	int goingNextCap;
	int levelSize[9][2];
	int totalClusters;
	int (*coarseTables)[4];
	REAL* dense; /* [block][96][96] row-major */
	REAL* carry; /* scratch: m_additionalHessian32 followed by the diagTable sums, [node][9] column-major each */
	REAL* inv;   /* [block][TRI] lower triangle row-major */
	REAL *R, *Z; /* [node][3] */
	int sorted;
} Oracle;

static uint32_t lanemask_lt(unsigned lane) { return (1u << lane) - 1u; }
static int ffs32(uint32_t x) { return x ? __builtin_ctz(x) + 1 : 0; }
static int pad32(int x) { return (x + 31) / 32 * 32; }

/* ---- Morton (SeMorton.h:75-101) ------------------------------------------ */
static uint64_t spread3(uint64_t b)
{
	b = (b | (b << 32)) & 0xFFFF00000000FFFFull;
	b = (b | (b << 16)) & 0x00FF0000FF0000FFull;
	b = (b | (b << 8)) & 0xF00F00F00F00F00Full;
	b = (b | (b << 4)) & 0x30C30C30C30C30C3ull;
	return (b | (b << 2)) & 0x9249249249249249ull;
}
/* Clamp(a,lo,hi) = Min(Max(lo,a),hi), Max(a,b)=(a>b)?a:b, Min(a,b)=(a<b)?a:b
 * (SeMath.h:100-103 with SePreDefine.h:37-38): a NaN falls through both
 * comparisons and comes out as `hi`. */
static float clamp_ref(float a, float lo, float hi)
{
	float t = (lo > a) ? lo : a;
	return (t < hi) ? t : hi;
}
static uint64_t axis_bits(float c)
{
	c = clamp_ref(c * 2097152.0f, 0.0f, 2097151.0f);
	return spread3((uint64_t)c);
}
uint64_t FN(morton_encode)(float x, float y, float z)
{
	return (axis_bits(x) << 2) + (axis_bits(y) << 1) + axis_bits(z);
}

/* ---- lifetime ------------------------------------------------------------ */
void* FN(create)(void) { return calloc(1, sizeof(Oracle)); }

static void free_prepare(Oracle* o)
{
	free(o->st); o->st = NULL;
	free(o->stIdx); o->stIdx = NULL;
	free(o->fineMask); o->fineMask = NULL;
	for (int l = 0; l < 8; ++l) { free(o->cst[l]); o->cst[l] = NULL; }
	free(o->goingNext); o->goingNext = NULL;
	free(o->coarseTables); o->coarseTables = NULL;
	free(o->dense); o->dense = NULL;
	free(o->carry); o->carry = NULL;
	free(o->inv); o->inv = NULL;
	free(o->R); o->R = NULL;
	free(o->Z); o->Z = NULL;
}

void FN(destroy)(void* h)
{
	Oracle* o = (Oracle*)h;
	if (!o) return;
	free_prepare(o);
	free(o->code); free(o->s2o); free(o->o2s);
	free(o->nbrStart); free(o->nbrIdx); free(o->nbrSrc);
	free(o->edges); free(o->faces);
	free(o);
}

void FN(set_option)(void* h, int which, int value)
{
	Oracle* o = (Oracle*)h;
	if (which == 0) o->prolongAllLevels = value;
}

/* ---- AllocatePrecoditioner (cpp:38-65) ----------------------------------- */
static int level_count(int nv) /* ComputeLevelNums, cpp:112-135 */
{
	int n = 1, sz = pad32(nv);
	while (sz > 32) { sz /= 32; ++n; sz = pad32(sz); }
	return n;
}

typedef struct { uint64_t code; int idx; } KeyIdx;
static int cmp_key(const void* a, const void* b)
{
	const KeyIdx *x = (const KeyIdx*)a, *y = (const KeyIdx*)b;
	if (x->code != y->code) return x->code < y->code ? -1 : 1;
	return (x->idx > y->idx) - (x->idx < y->idx);
}

int FN(allocate)(void* h, int nv, int ne, int nf, const float* positions, const int* edges, const int* faces,
	const int* nbrStarts, const int* nbrIdxIn)
{
	Oracle* o = (Oracle*)h;
	if (o->sorted) return 0; /* Q1: m_frameIndex sticks at 1, reorder runs once per object (cpp:44-64) */
	o->nv = nv; o->ne = ne; o->nf = nf; o->nnz = nbrStarts[nv];
	o->numLevel = level_count(nv);
	if (o->numLevel > 5) return -2; /* Int4 coarse table */

	/* ComputeAABB cpp:201-211; SeAabbSimd.h:76-79 is _mm_min_ps/_mm_max_ps on all four lanes */
	for (int c = 0; c < 4; ++c) { o->lo[c] = 3.402823466e+38f; o->hi[c] = -3.402823466e+38f; }
	o->lo[3] = 0.f; o->hi[3] = 0.f;
	for (int v = 0; v < nv; ++v)
		for (int c = 0; c < 4; ++c)
		{
			float p = positions[4 * v + c];
			/* minps(a,b) = a<b ? a : b with a = current bound */
			o->lo[c] = (o->lo[c] < p) ? o->lo[c] : p;
			o->hi[c] = (o->hi[c] > p) ? o->hi[c] : p;
		}

	/* FillSortingData cpp:219-235 */
	o->code = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)nv);
	KeyIdx* keys = (KeyIdx*)malloc(sizeof(KeyIdx) * (size_t)nv);
	for (int v = 0; v < nv; ++v)
	{
		float t[3];
		for (int c = 0; c < 3; ++c)
		{
			float ext = o->hi[c] - o->lo[c];
			t[c] = (positions[4 * v + c] - o->lo[c]) / ext;
		}
		o->code[v] = FN(morton_encode)(t[0], t[1], t[2]);
		keys[v].code = o->code[v];
		keys[v].idx = v;
	}
	/* DoingSort cpp:238-243 (ties: ascending original index) */
	qsort(keys, (size_t)nv, sizeof(KeyIdx), cmp_key);
	o->s2o = (int*)malloc(sizeof(int) * (size_t)nv);
	o->o2s = (int*)malloc(sizeof(int) * (size_t)nv);
	for (int v = 0; v < nv; ++v) { o->s2o[v] = keys[v].idx; o->o2s[keys[v].idx] = v; } /* cpp:245-255 */
	free(keys);

	/* MapHessianTable cpp:258-285, as a CSR in sorted space instead of the SoA table */
	o->nbrStart = (int*)malloc(sizeof(int) * ((size_t)nv + 1));
	o->nbrIdx = (int*)malloc(sizeof(int) * (size_t)(o->nnz > 0 ? o->nnz : 1));
	o->nbrSrc = (int*)malloc(sizeof(int) * (size_t)(o->nnz > 0 ? o->nnz : 1));
	int pos = 0;
	for (int v = 0; v < nv; ++v)
	{
		int ov = o->s2o[v];
		o->nbrStart[v] = pos;
		for (int k = nbrStarts[ov]; k < nbrStarts[ov + 1]; ++k)
		{
			o->nbrIdx[pos] = o->o2s[nbrIdxIn[k]];
			o->nbrSrc[pos] = k;
			++pos;
		}
	}
	o->nbrStart[nv] = pos;

	o->edges = (int*)malloc(sizeof(int) * 4 * (size_t)(ne > 0 ? ne : 1));
	o->faces = (int*)malloc(sizeof(int) * 4 * (size_t)(nf > 0 ? nf : 1));
	if (ne > 0) memcpy(o->edges, edges, sizeof(int) * 4 * (size_t)ne);
	if (nf > 0) memcpy(o->faces, faces, sizeof(int) * 4 * (size_t)nf);
	o->sorted = 1;
	return 0;
}

/* ---- PrepareCollisionStencils (cpp:304-413) ------------------------------ */
static float rd_f(const unsigned char* rec, int off) { float f; memcpy(&f, rec + off, 4); return f; }
static int rd_i(const unsigned char* rec, int off) { int i; memcpy(&i, rec + off, 4); return i; }

static void build_stencils(Oracle* o, const void* efv, const void* eev, const void* vfv,
	unsigned efN, unsigned eeN, unsigned vfN)
{
	const unsigned char *ef = (const unsigned char*)efv, *ee = (const unsigned char*)eev, *vf = (const unsigned char*)vfv;
	long total = (long)efN + eeN + vfN;
	long cap = (long)o->nv * 32; /* cpp:187-188 */
	if (total > cap) total = cap;   /* cpp:312-316 */
	o->st = (Stencil*)calloc((size_t)(total > 0 ? total : 1), sizeof(Stencil));
	o->stIdx = (int(*)[5])calloc((size_t)(total > 0 ? total : 1), sizeof(int[5]));
	int n = 0;
	for (long i = 0; i < total; ++i)
	{
		Stencil s;
		memset(&s, 0, sizeof s);
		if (i < (long)efN)
		{
			const unsigned char* p = ef + 48 * i; /* EfSet: eId@0 fId@4 stiff@8 bary@12 normal@32 */
			int e = rd_i(p, 0), f = rd_i(p, 4);
			if (e < 0 || f < 0) continue;
			float b0 = rd_f(p, 12), b1 = rd_f(p, 16), b2 = rd_f(p, 20);
			s.n = 5; s.nFirst = 2;
			s.index[0] = o->edges[4 * e]; s.index[1] = o->edges[4 * e + 1];
			s.index[2] = o->faces[4 * f]; s.index[3] = o->faces[4 * f + 1]; s.index[4] = o->faces[4 * f + 2];
			s.weight[0] = b0; s.weight[1] = 1.f - b0;
			s.weight[2] = -b1; s.weight[3] = -b2; s.weight[4] = -(1.f - b1 - b2);
			for (int c = 0; c < 4; ++c) s.dir[c] = rd_f(p, 32 + 4 * c);
			s.stiff = rd_f(p, 8);
		}
		else if (i < (long)efN + eeN)
		{
			const unsigned char* p = ee + 48 * i; /* Q2: global index (cpp:357); EeSet bary@16 */
			int e0 = rd_i(p, 0), e1 = rd_i(p, 4);
			if (e1 < 0 || e0 < 0) continue;
			float b0 = rd_f(p, 16), b1 = rd_f(p, 20);
			s.n = 4; s.nFirst = 2;
			s.index[0] = o->edges[4 * e0]; s.index[1] = o->edges[4 * e0 + 1];
			s.index[2] = o->edges[4 * e1]; s.index[3] = o->edges[4 * e1 + 1];
			s.weight[0] = b0; s.weight[1] = 1.f - b0;
			s.weight[2] = -b1; s.weight[3] = -(1.f - b1);
			for (int c = 0; c < 4; ++c) s.dir[c] = rd_f(p, 32 + 4 * c);
			s.stiff = rd_f(p, 8);
		}
		else
		{
			const unsigned char* p = vf + 48 * i; /* Q2 global index (cpp:383); Q3: m_bary[2] = float at byte 24 (cpp:399) */
			int v = rd_i(p, 0), f = rd_i(p, 4);
			if (v < 0 || f < 0) continue;
			float b0 = rd_f(p, 16), b1 = rd_f(p, 20), b2 = rd_f(p, 24);
			s.n = 4; s.nFirst = 3;
			s.index[0] = o->faces[4 * f]; s.index[1] = o->faces[4 * f + 1]; s.index[2] = o->faces[4 * f + 2];
			s.index[3] = v;
			s.weight[0] = -b0; s.weight[1] = -b1; s.weight[2] = -(1.f - b2); s.weight[3] = 1.f;
			for (int c = 0; c < 4; ++c) s.dir[c] = rd_f(p, 32 + 4 * c);
			s.stiff = rd_f(p, 8);
		}
		o->st[n] = s;
		for (int k = 0; k < s.n; ++k) o->stIdx[n][k] = o->o2s[s.index[k]]; /* cpp:287-302 */
		++n;
	}
	o->nStencil = n;
}

/* ---- clustering (ReorderRealtime, cpp:415-445) --------------------------- */

/* BuildCollisionConnection cpp:514-563 */
static void collision_connect(Oracle* o, uint32_t* mask, const int* coarse)
{
	for (int i = 0; i < o->nStencil; ++i)
	{
		const Stencil* s = &o->st[i];
		unsigned id[5];
		uint32_t m[5] = { 0, 0, 0, 0, 0 };
		for (int k = 0; k < s->n; ++k) id[k] = (unsigned)(coarse ? coarse[o->stIdx[i][k]] : o->stIdx[i][k]);
		for (int a = 0; a < s->n; ++a)
			for (int b = a + 1; b < s->n; ++b)
			{
				if (id[a] == id[b]) continue;
				if (id[a] / BANK != id[b] / BANK) continue;
				if (a < s->nFirst && b >= s->nFirst)
				{
					m[a] |= 1u << (id[b] % BANK);
					m[b] |= 1u << (id[a] % BANK);
				}
			}
		for (int k = 0; k < s->n; ++k)
			if (m[k]) mask[id[k]] |= m[k];
	}
}

/* In-bank transitive closure by bit flood-fill (cpp:596-614, 926-944) for
 * count nodes whose masks are in mask[]; writes the closures back. */
static void close_components(uint32_t* mask, int count)
{
	for (int base = 0; base < count; base += BANK)
	{
		uint32_t cache[BANK];
		int n = count - base < BANK ? count - base : BANK;
		for (int l = 0; l < BANK; ++l) cache[l] = l < n ? mask[base + l] : (1u << l);
		for (int l = 0; l < n; ++l)
		{
			uint32_t m = cache[l], seen = 1u << l;
			for (;;)
			{
				uint32_t todo = seen ^ m;
				if (!todo) break;
				int nx = ffs32(todo) - 1;
				seen |= 1u << nx;
				m |= cache[nx];
			}
			mask[base + l] = m;
		}
	}
}

/* Number the components: id of a node = (#elected nodes in earlier banks) +
 * rank of its component's lowest lane among the bank's elected lanes
 * (cpp:663-737 for level 0, cpp:1007-1069 for level >= 1).  Returns the count. */
static int number_components(const uint32_t* mask, int count, int* idOut)
{
	int running = 0;
	for (int base = 0; base < count; base += BANK)
	{
		int n = count - base < BANK ? count - base : BANK;
		uint32_t elected = 0;
		for (int l = 0; l < n; ++l)
			if ((mask[base + l] & lanemask_lt((unsigned)l)) == 0) elected |= 1u << l;
		for (int l = 0; l < n; ++l)
		{
			int rep = ffs32(mask[base + l]) - 1;
			idOut[base + l] = running + __builtin_popcount(elected & lanemask_lt((unsigned)rep));
		}
		running += __builtin_popcount(elected);
	}
	return running;
}

static void build_hierarchy(Oracle* o)
{
	const int nv = o->nv, L = o->numLevel;
	const int nVC = pad32(nv);
	memset(o->levelSize, 0, sizeof o->levelSize);
	o->fineMask = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)nv);
	for (int l = 0; l < L; ++l) o->cst[l] = (int*)malloc(sizeof(int) * (size_t)nv);
	/* generous first guess; grown below if the hierarchy is larger (Q6) */
	o->goingNextCap = nVC + nVC / 8 + 4096;
	o->goingNext = (int*)calloc((size_t)o->goingNextCap, sizeof(int));

	/* BuildConnectMaskL0 cpp:447-511 */
	for (int v = 0; v < nv; ++v)
	{
		uint32_t m = 1u << (v % BANK);
		for (int k = o->nbrStart[v]; k < o->nbrStart[v + 1]; ++k)
			if (o->nbrIdx[k] / BANK == v / BANK) m |= 1u << (o->nbrIdx[k] % BANK);
		o->fineMask[v] = m;
	}
	collision_connect(o, o->fineMask, NULL);
	close_components(o->fineMask, nv);               /* PreparePrefixSumL0 cpp:565-628 */
	int n1 = number_components(o->fineMask, nv, o->cst[0]); /* BuildLevel1 cpp:630-740 */
	for (int v = 0; v < nv; ++v) o->goingNext[v] = o->cst[0][v] + nVC;
	o->levelSize[1][0] = n1;
	o->levelSize[1][1] = nVC;

	/* a level never has more nodes than the one below; masks are cleared in whole banks (pad32(cnt) <= nVC) */
	uint32_t* nextMask = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)nVC + 32));
	int* nextId = (int*)malloc(sizeof(int) * ((size_t)nVC + 32));
	for (int level = 1; level < L; ++level)
	{
		const int cnt = o->levelSize[level][0], begin = o->levelSize[level][1];
		const int* coarse = o->cst[level - 1];
		int need = begin + pad32(cnt) + 64;
		if (need > o->goingNextCap)
		{
			int old = o->goingNextCap;
			o->goingNextCap = need * 2;
			o->goingNext = (int*)realloc(o->goingNext, sizeof(int) * (size_t)o->goingNextCap);
			memset(o->goingNext + old, 0, sizeof(int) * (size_t)(o->goingNextCap - old));
		}
		memset(nextMask, 0, sizeof(uint32_t) * (size_t)(cnt > 0 ? pad32(cnt) : 32));
		/* BuildConnectMaskLx cpp:743-871: the per-component OR + elected-lane atomicOr
		 * is an OR over every fine vertex of the coarse node */
		for (int v = 0; v < nv; ++v)
		{
			unsigned cv = (unsigned)coarse[v];
			for (int k = o->nbrStart[v]; k < o->nbrStart[v + 1]; ++k)
			{
				unsigned cu = (unsigned)coarse[o->nbrIdx[k]];
				if (cv / BANK == cu / BANK) nextMask[cv] |= 1u << (cu % BANK);
			}
		}
		collision_connect(o, nextMask, coarse);
		for (int c = 0; c < cnt; ++c) nextMask[c] |= 1u << (c % BANK); /* NextLevelCluster cpp:898-903 */
		close_components(nextMask, cnt);                               /* cpp:917-954 */
		int nNext = number_components(nextMask, cnt, nextId);          /* PrefixSumLx cpp:963-1072 (scan done right, Q5) */
		for (int c = 0; c < cnt; ++c) o->goingNext[begin + c] = nextId[c] + begin + pad32(cnt);
		o->levelSize[level + 1][0] = nNext;
		o->levelSize[level + 1][1] = begin + pad32(cnt);
		for (int v = 0; v < nv; ++v) o->cst[level][v] = nextId[coarse[v]]; /* ComputeNextLevel cpp:1074-1084 */
	}
	free(nextMask);
	free(nextId);
	o->totalClusters = o->levelSize[L][1]; /* TotalNodes cpp:1086-1090 */

	/* AggregationKernel cpp:1092-1162 */
	o->coarseTables = (int(*)[4])calloc((size_t)nv, sizeof(int[4]));
	for (int v = 0; v < nv; ++v)
	{
		int cur = v;
		for (int l = 0; l < L - 1; ++l) { cur = o->goingNext[cur]; o->coarseTables[v][l] = cur; }
	}
}

/* ---- assembly (cpp:1164-1345) -------------------------------------------- */
/* m is a column-major 3x3 (SeMatrix.h:681-682); dense block is row-major 96x96 */
static void dense_add(Oracle* o, unsigned rowNode, unsigned colNode, const REAL m[9])
{
	REAL* D = o->dense + (size_t)(rowNode / BANK) * DOF * DOF;
	int r0 = 3 * (int)(rowNode % BANK), c0 = 3 * (int)(colNode % BANK);
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j) D[(r0 + i) * DOF + c0 + j] += m[3 * j + i];
}
static void add9(REAL* dst, const REAL* src, REAL scale)
{
	for (int k = 0; k < 9; ++k) dst[k] += scale * src[k];
}

static void assemble(Oracle* o, const float* diag, const float* offdiag)
{
	const int nv = o->nv, L = o->numLevel, total = o->totalClusters;
	const int nVC = pad32(nv);
	const size_t nBlocks = (size_t)total / BANK;
	o->dense = (REAL*)calloc(nBlocks * DOF * DOF, sizeof(REAL));
	o->carry = (REAL*)calloc(((size_t)total + 1) * 9 * 2, sizeof(REAL));
	REAL* extra = o->carry;                          /* m_additionalHessian32 (h:108) */
	REAL* table = o->carry + ((size_t)total + 1) * 9; /* the per-level diagTable maps (cpp:1257), keyed by node id */

	/* PrepareCollisionHessian cpp:1201-1227 with AdditionalSchwarzHessian2 cpp:1164-1199 */
	for (int i = 0; i < o->nStencil; ++i)
	{
		const Stencil* s = &o->st[i];
		REAL H[9];
		for (int a = 0; a < 3; ++a)
			for (int b = 0; b < 3; ++b) H[3 * b + a] = (REAL)s->dir[a] * ((REAL)s->dir[b] * (REAL)s->stiff);
		for (int k = 0; k < s->n; ++k) add9(extra + 9 * (size_t)o->stIdx[i][k], H, (REAL)s->weight[k] * (REAL)s->weight[k]);
		for (int a = 0; a < s->n; ++a)
			for (int b = a + 1; b < s->n; ++b)
			{
				REAL Hp[9];
				REAL w = (REAL)s->weight[a] * (REAL)s->weight[b];
				for (int k = 0; k < 9; ++k) Hp[k] = w * H[k];
				unsigned my = (unsigned)o->stIdx[i][a], ot = (unsigned)o->stIdx[i][b];
				int level = 0;
				while (my / BANK != ot / BANK && level < L) { my = o->goingNext[my]; ot = o->goingNext[ot]; ++level; }
				if (level >= L) continue;
				dense_add(o, my, ot, Hp);
				dense_add(o, ot, my, Hp);
				if (level < L - 1)
				{
					my = o->goingNext[my]; ot = o->goingNext[ot];
					if (my == ot) add9(extra + 9 * (size_t)my, Hp, (REAL)2);
					else { add9(extra + 9 * (size_t)my, Hp, (REAL)1); add9(extra + 9 * (size_t)ot, Hp, (REAL)1); }
				}
			}
	}

	/* PrepareHessian part A, cpp:1238-1252: a coarse node's collision terms go
	 * onto its own diagonal block and every ancestor's */
	for (int c = nVC; c < total; ++c)
		for (int n = c; n < total; n = o->goingNext[n]) dense_add(o, (unsigned)n, (unsigned)n, extra + 9 * (size_t)c);

	/* part B, cpp:1254-1324, in the one-thread order of the reference */
	for (int v = 0; v < nv; ++v)
	{
		const int ov = o->s2o[v];
		REAL D[9];
		for (int k = 0; k < 9; ++k) D[k] = (REAL)diag[9 * (size_t)ov + k] + extra[9 * (size_t)v + k];
		dense_add(o, (unsigned)v, (unsigned)v, D); /* cpp:1270-1271 */
		for (int e = o->nbrStart[v]; e < o->nbrStart[v + 1]; ++e)
		{
			REAL M[9];
			for (int k = 0; k < 9; ++k) M[k] = (REAL)offdiag[9 * (size_t)o->nbrSrc[e] + k];
			unsigned my = (unsigned)v, ot = (unsigned)o->nbrIdx[e];
			int level = 0;
			while (my / BANK != ot / BANK && level < L) { ++level; my = o->goingNext[my]; ot = o->goingNext[ot]; }
			if (level >= L) continue;                 /* cpp:1288-1291 */
			dense_add(o, my, ot, M);                   /* cpp:1292-1295 */
			if (level == 0) add9(D, M, (REAL)1);       /* cpp:1297-1298 */
			else if (level + 1 < L) add9(table + 9 * (size_t)o->goingNext[my], M, (REAL)1); /* cpp:1299-1307 */
		}
		if (L > 1)
		{
			int n1 = o->goingNext[v];
			dense_add(o, (unsigned)n1, (unsigned)n1, D);                      /* cpp:1311-1312 */
			if (L > 2) add9(table + 9 * (size_t)o->goingNext[n1], D, (REAL)1); /* cpp:1313-1321 */
		}
	}
	/* table flush, cpp:1326-1343 (map iteration order there is unspecified; ascending node id here) */
	for (int level = 2; level < L; ++level)
	{
		const int begin = o->levelSize[level][1], cnt = o->levelSize[level][0];
		for (int c = begin; c < begin + cnt; ++c)
		{
			dense_add(o, (unsigned)c, (unsigned)c, table + 9 * (size_t)c);
			if (level + 1 < L) add9(table + 9 * (size_t)o->goingNext[c], table + 9 * (size_t)c, (REAL)1);
		}
	}
}

/* ---- LDLtInverse512 (cpp:1347-1546) -------------------------------------- */
#if defined(ORACLE_DOUBLE)
#define FMA(a, b, c) fma((a), (b), (c))
#else
#define FMA(a, b, c) fmaf((a), (b), (c))
#endif

static void invert_block(const REAL* dense, REAL* outTri)
{
	static __thread REAL A[DOF][DOF];
	REAL dinv[DOF];
	memcpy(A, dense, sizeof(REAL) * DOF * DOF);
	for (int n = 0; n < BANK; ++n) /* padding nodes -> identity, cpp:1365-1368 */
		if (A[3 * n][3 * n] == (REAL)0)
			for (int i = 0; i < 3; ++i)
				for (int j = 0; j < 3; ++j) A[3 * n + i][3 * n + j] = (i == j) ? (REAL)1 : (REAL)0;
	/* un-pivoted row elimination on full rows; the multipliers left below the
	 * diagonal accumulate into E = L^-1 (cpp:1395-1415) */
	for (int x = 0; x < DOF; ++x)
	{
		const REAL d = A[x][x];
		for (int y = x + 1; y < DOF; ++y)
		{
			if (A[y][x] == (REAL)0) continue;
			const REAL r = -A[y][x] / d;
			for (int c = 0; c < DOF; ++c) A[y][c] = FMA(r, A[x][c], A[y][c]);
			A[y][x] = r;
		}
	}
	for (int y = 0; y < DOF; ++y) /* cpp:1419-1433 */
	{
		dinv[y] = (REAL)1 / A[y][y];
		A[y][y] = (REAL)1;
	}
	/* inv = E^T D^-1 E, lower triangle; rows summed from 95 downwards (cpp:1437-1495) */
	for (int r = 0; r < DOF; ++r)
		for (int c = 0; c <= r; ++c)
		{
			REAL acc = (REAL)0;
			for (int p = DOF - 1; p >= r; --p)
			{
				REAL e = (p == r) ? A[p][c] : A[p][c] * A[p][r];
				acc = FMA(dinv[p], e, acc);
			}
			outTri[r * (r + 1) / 2 + c] = acc;
		}
}

/* ---- PreparePreconditioner (cpp:67-98) ----------------------------------- */
int FN(prepare)(void* h, const float* diag, const float* offdiag, const int* ranges,
	const void* ef, const void* ee, const void* vf, unsigned efN, unsigned eeN, unsigned vfN)
{
	Oracle* o = (Oracle*)h;
	(void)ranges; /* equals the adjacency starts (cpp:1276 with cpp:272-282) */
	free_prepare(o);
	build_stencils(o, ef, ee, vf, efN, eeN, vfN);
	build_hierarchy(o);
	assemble(o, diag, offdiag);
	const int nBlocks = o->totalClusters / BANK;
	o->inv = (REAL*)malloc(sizeof(REAL) * (size_t)nBlocks * TRI);
#pragma omp parallel for schedule(dynamic, 16)
	for (int b = 0; b < nBlocks; ++b) invert_block(o->dense + (size_t)b * DOF * DOF, o->inv + (size_t)b * TRI);
	o->R = (REAL*)calloc((size_t)o->totalClusters * 3, sizeof(REAL));
	o->Z = (REAL*)calloc((size_t)o->totalClusters * 3, sizeof(REAL));
	return 0;
}

/* ---- Preconditioning (cpp:100-110) --------------------------------------- */
int FN(apply)(void* h, float* z, const float* residual)
{
	Oracle* o = (Oracle*)h;
	const int nv = o->nv, L = o->numLevel, total = o->totalClusters;
	REAL *R = o->R, *Z = o->Z;
	memset(R, 0, sizeof(REAL) * 3 * (size_t)total);
	/* BuildResidualHierarchy cpp:1548-1598 */
	for (int v = 0; v < nv; ++v)
	{
		const float* r = residual + 4 * (size_t)o->s2o[v];
		for (int c = 0; c < 3; ++c) R[3 * (size_t)v + c] = (REAL)r[c];
		if (L > 1)
			for (int c = 0; c < 3; ++c) R[3 * (size_t)o->goingNext[v] + c] += (REAL)r[c];
	}
	for (int level = 1; level + 1 < L; ++level)
	{
		const int begin = o->levelSize[level][1], cnt = o->levelSize[level][0];
		for (int n = begin; n < begin + cnt; ++n)
			for (int c = 0; c < 3; ++c) R[3 * (size_t)o->goingNext[n] + c] += R[3 * (size_t)n + c];
	}
	/* SchwarzLocalXSym cpp:1600-1696: Z = blockdiag(inv) R on every active block */
	const int nBlocks = total / BANK;
#pragma omp parallel for schedule(static)
	for (int b = 0; b < nBlocks; ++b)
	{
		const REAL* T = o->inv + (size_t)b * TRI;
		const REAL* x = R + (size_t)b * DOF;
		REAL y[DOF];
		for (int r = 0; r < DOF; ++r) y[r] = T[r * (r + 1) / 2 + r] * x[r];
		for (int r = 1; r < DOF; ++r)
		{
			const REAL* row = T + r * (r + 1) / 2;
			REAL acc = (REAL)0;
			for (int c = 0; c < r; ++c)
			{
				acc += row[c] * x[c];
				y[c] += row[c] * x[r];
			}
			y[r] += acc;
		}
		memcpy(Z + (size_t)b * DOF, y, sizeof y);
	}
	/* CollectFinalZ cpp:1698-1719 (Q4: at most levels 1..3 are prolonged) */
	const int top = o->prolongAllLevels ? L : (L < 4 ? L : 4);
	for (int v = 0; v < nv; ++v)
	{
		REAL acc[3] = { Z[3 * (size_t)v], Z[3 * (size_t)v + 1], Z[3 * (size_t)v + 2] };
		for (int l = 1; l < top; ++l)
		{
			const int n = o->coarseTables[v][l - 1];
			for (int c = 0; c < 3; ++c) acc[c] += Z[3 * (size_t)n + c];
		}
		float* out = z + 4 * (size_t)o->s2o[v];
		out[0] = (float)acc[0]; out[1] = (float)acc[1]; out[2] = (float)acc[2]; out[3] = 0.f;
	}
	return 0;
}

/* ---- introspection -------------------------------------------------------- */
int FN(num_level)(void* h) { return ((Oracle*)h)->numLevel; }
int FN(total_clusters)(void* h) { return ((Oracle*)h)->totalClusters; }
int FN(stencil_num)(void* h) { return ((Oracle*)h)->nStencil; }
void FN(get_aabb)(void* h, float* lo, float* hi) { Oracle* o = (Oracle*)h; memcpy(lo, o->lo, 16); memcpy(hi, o->hi, 16); }
void FN(get_level_size)(void* h, int* out)
{
	Oracle* o = (Oracle*)h;
	for (int l = 0; l <= o->numLevel; ++l) { out[2 * l] = o->levelSize[l][0]; out[2 * l + 1] = o->levelSize[l][1]; }
}
void FN(get_morton)(void* h, uint64_t* out) { Oracle* o = (Oracle*)h; memcpy(out, o->code, 8 * (size_t)o->nv); }
void FN(get_sorted_get_original)(void* h, int* out) { Oracle* o = (Oracle*)h; memcpy(out, o->s2o, 4 * (size_t)o->nv); }
void FN(get_original_get_sorted)(void* h, int* out) { Oracle* o = (Oracle*)h; memcpy(out, o->o2s, 4 * (size_t)o->nv); }
void FN(get_going_next)(void* h, int* out, int count) { Oracle* o = (Oracle*)h; memcpy(out, o->goingNext, 4 * (size_t)count); }
void FN(get_coarse_tables)(void* h, int* out) { Oracle* o = (Oracle*)h; memcpy(out, o->coarseTables, 16 * (size_t)o->nv); }
void FN(get_coarse_space_table)(void* h, int level, int* out) { Oracle* o = (Oracle*)h; memcpy(out, o->cst[level], 4 * (size_t)o->nv); }
void FN(get_fine_connect_mask)(void* h, uint32_t* out) { Oracle* o = (Oracle*)h; memcpy(out, o->fineMask, 4 * (size_t)o->nv); }
void FN(get_sorted_adjacency)(void* h, int* starts, int* idx)
{
	Oracle* o = (Oracle*)h;
	memcpy(starts, o->nbrStart, 4 * ((size_t)o->nv + 1));
	memcpy(idx, o->nbrIdx, 4 * (size_t)o->nnz);
}
void FN(get_stencils)(void* h, void* stencils80, int* mapped)
{
	Oracle* o = (Oracle*)h;
	unsigned char* dst = (unsigned char*)stencils80;
	for (int i = 0; i < o->nStencil; ++i)
	{
		const Stencil* s = &o->st[i];
		unsigned char* p = dst + 80 * (size_t)i; /* n@0 nFirst@4 index@8 weight@28 stiff@48 direction@64 */
		memset(p, 0, 80);
		memcpy(p, &s->n, 4); memcpy(p + 4, &s->nFirst, 4);
		memcpy(p + 8, s->index, 20); memcpy(p + 28, s->weight, 20);
		memcpy(p + 48, &s->stiff, 4); memcpy(p + 64, s->dir, 16);
		memcpy(mapped + 5 * (size_t)i, o->stIdx[i], 20);
	}
}
void FN(get_dense_hessian)(void* h, int block, REAL* out)
{
	Oracle* o = (Oracle*)h;
	memcpy(out, o->dense + (size_t)block * DOF * DOF, sizeof(REAL) * DOF * DOF);
}
void FN(get_dense_inverse)(void* h, int block, REAL* out)
{
	Oracle* o = (Oracle*)h;
	const REAL* T = o->inv + (size_t)block * TRI;
	for (int r = 0; r < DOF; ++r)
		for (int c = 0; c <= r; ++c) { out[r * DOF + c] = T[r * (r + 1) / 2 + c]; out[c * DOF + r] = out[r * DOF + c]; }
}
/* all packed lower triangles, [nBlocks][4656] */
void FN(get_packed_inverses)(void* h, REAL* out)
{
	Oracle* o = (Oracle*)h;
	memcpy(out, o->inv, sizeof(REAL) * (size_t)(o->totalClusters / BANK) * TRI);
}
void FN(get_mapped_r)(void* h, REAL* out) { Oracle* o = (Oracle*)h; memcpy(out, o->R, sizeof(REAL) * 3 * (size_t)o->totalClusters); }
void FN(get_mapped_z)(void* h, REAL* out) { Oracle* o = (Oracle*)h; memcpy(out, o->Z, sizeof(REAL) * 3 * (size_t)o->totalClusters); }
int FN(sizeof_real)(void) { return (int)sizeof(REAL); }
