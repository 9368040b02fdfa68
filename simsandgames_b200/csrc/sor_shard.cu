// Slab-sharded red-black SOR projection: one process per GPU, one slab of grid rows per rank, nearest-neighbour
// NCCL exchanges over NVLink (stage "F9").
//
// The grid is cut along the SLOW storage axis (rows are contiguous, so a block of halo rows is one contiguous
// message).  Rank r owns cell rows [j0, j1) and keeps H halo rows on each side.  Communication-avoiding deep
// halo: after an exchange every local row is an exact copy of its owner's row; a colour pass then updates all
// rows whose inputs are still exact -- own rows plus a halo band that shrinks by one row per pass -- so K = H-1
// passes run back to back with no communication, the last one still covering own rows +-1 (which completes
// the face row two slabs share).  The redundant halo cells perform exactly the updates their owner performs,
// on identical inputs, so the result is bit-identical to the single-GPU red-black solve
// (tests/test_gpu_multi.py) at 2*iters/K exchanges per solve instead of 2*iters (H=9: 50 instead of 400 for the
// 200-sweep config), each carrying H rows of the two velocity fields.
// The lexicographic wavefront is not sharded: its critical path crosses all slabs.
#include <algorithm>
#include <cstdlib>
#include "common.cuh"
#define FG_NCCL_SHIM_IMPL
#include "nccl_shim.cuh"

namespace fg {

// one colour of one SOR sweep over local rows [b_lo, b_hi]; `shift` = parity of the global index of local row 0
template <bool XFAST>
__global__ void __launch_bounds__(256) sor_rb_rows_kernel(SorGrid g, int colour, int shift, int b_lo, int b_hi, float cp,
                                                          float omega, bool drift) {
	int b = b_lo + blockIdx.y * blockDim.y + threadIdx.y;
	if (b > b_hi) return;
	int k = blockIdx.x * blockDim.x + threadIdx.x;
	int a = 2 * k + ((b + shift + colour) & 1);
	if (a < 1 || a > g.nf - 2) return;
	const int nf = g.nf, c = a + nf * b;
	if (g.cell_type[c] != FLUID_CELL) return;
	float sfm = g.s[c - 1], sfp = g.s[c + 1], ssm = g.s[c - nf], ssp = g.s[c + nf];
	float ssum = XFAST ? ((sfm + sfp) + ssm) + ssp : ((ssm + ssp) + sfm) + sfp;
	if (ssum == 0.f) return;
	float vfc = g.vel_f[c], vfp = g.vel_f[c + 1], vsc = g.vel_s[c], vsp = g.vel_s[c + nf];
	float div = XFAST ? ((vfp - vfc) + vsp) - vsc : ((vsp - vsc) + vfp) - vfc;
	float rest = (drift && g.rest) ? *g.rest : 0.f;
	if (drift && rest > 0.f) {
		float compression = g.density[c] - rest;
		if (compression > 0.f) div -= compression;
	}
	float pv = -div / ssum;
	pv *= omega;
	g.p[c] += cp * pv;
	g.vel_f[c] = vfc - sfm * pv;
	g.vel_f[c + 1] = vfp + sfp * pv;
	g.vel_s[c] = vsc - ssm * pv;
	g.vel_s[c + nf] = vsp + ssp * pv;
}

// one colour pass over global rows [lo, hi] of a grid view (used by the sharded FLIP step as well)
void sor_rb_rows_launch(const SorGrid& g, int colour, int shift, int lo, int hi, float cp, float omega, bool drift, cudaStream_t st) {
	if (hi < lo) return;
	dim3 block(64, 4);
	dim3 grid(cdiv((size_t)(g.nf + 1) / 2 + 1, block.x), cdiv((size_t)(hi - lo + 1), block.y));
	if (g.x_fast) sor_rb_rows_kernel<true><<<grid, block, 0, st>>>(g, colour, shift, lo, hi, cp, omega, drift);
	else sor_rb_rows_kernel<false><<<grid, block, 0, st>>>(g, colour, shift, lo, hi, cp, omega, drift);
}

}  // namespace fg

using namespace fg;

struct SorShard {
	int nf = 0, ns = 0, rank = 0, nranks = 1, j0 = 0, j1 = 0, rows = 0, device = 0;
	int halo = 9;  // H halo rows per side -> K = H-1 colour passes between exchanges
	bool x_fast = true;
	float *vel_f = nullptr, *vel_s = nullptr, *p = nullptr, *s = nullptr, *dens = nullptr, *d_rest = nullptr;
	float *alt_f = nullptr, *alt_s = nullptr, *alt_p = nullptr;  // ping-pong partners (tiled red-black rounds, lazy)
	SorWorkspace* ws = nullptr;
	bool s_binary = true;  // every uploaded s value was 0 or 1 (the tiled kernel's 1-byte cell code needs that)
	int* cell_type = nullptr;
	ncclComm_t comm = nullptr;
	cudaStream_t st = nullptr;
	LaunchCounter lc;
	long long exchanges = 0;
};

namespace {
struct SDeviceGuard {
	int prev = -1;
	explicit SDeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); }
	~SDeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
void* shard_field(SorShard* h, int field, size_t* elem) {
	*elem = 4;
	switch (field) {
		case SORSHARD_VEL_FAST: return h->vel_f;
		case SORSHARD_VEL_SLOW: return h->vel_s;
		case SORSHARD_P: return h->p;
		case SORSHARD_S: return h->s;
		case SORSHARD_DENSITY: return h->dens;
		case SORSHARD_CELL_TYPE: return h->cell_type;
	}
	return nullptr;
}
}  // namespace

extern "C" {

int flipgpu_nccl_unique_id(void* out128) {
	FG_CHECK(out128, FLIPGPU_EINVAL, "null argument");
	FG_TRY(load_nccl());
	ncclUniqueId id;
	FG_NCCL(g_nccl.GetUniqueId(&id));
	static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
	memcpy(out128, &id, sizeof(id));
	return FLIPGPU_OK;
}

void sorshard_slab(int ns, int rank, int nranks, int* j0, int* j1) {
	*j0 = (int)((long long)ns * rank / nranks);
	*j1 = (int)((long long)ns * (rank + 1) / nranks);
}

int sorshard_create(int nf, int ns, int x_fast, int rank, int nranks, const void* unique_id128, int device, SorShard** out) {
	FG_CHECK(out && nf >= 3 && ns >= 3 && nranks >= 1 && rank >= 0 && rank < nranks, FLIPGPU_EINVAL, "sorshard_create: bad argument");
	int halo = 9;
	if (const char* e = getenv("FLIPGPU_SHARD_HALO")) halo = std::max(2, atoi(e));
	if (nranks == 1) halo = 1;
	FG_CHECK(nranks == 1 || ns / nranks >= halo, FLIPGPU_EINVAL, "sorshard_create: %d rows over %d ranks leaves slabs thinner than the %d-row halo", ns, nranks, halo);
	FG_CHECK(nranks == 1 || unique_id128, FLIPGPU_EINVAL, "sorshard_create: a NCCL unique id is required for %d ranks", nranks);
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
		fg::set_error("sorshard_create: no CUDA device (this library has no CPU fallback)");
		return FLIPGPU_ENODEV;
	}
	if (device < 0) FG_CUDA(cudaGetDevice(&device));
	FG_CHECK(device < ndev, FLIPGPU_EINVAL, "sorshard_create: device %d of %d", device, ndev);
	SorShard* h = new SorShard;
	h->nf = nf; h->ns = ns; h->rank = rank; h->nranks = nranks; h->device = device; h->x_fast = x_fast != 0;
	sorshard_slab(ns, rank, nranks, &h->j0, &h->j1);
	h->halo = halo;
	h->rows = h->j1 - h->j0 + 2 * halo;  // own rows + H halo rows on each side
	SDeviceGuard guard(device);
	FG_CUDA(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
	size_t n = (size_t)nf * h->rows;
	float** fs[] = {&h->vel_f, &h->vel_s, &h->p, &h->s, &h->dens};
	for (float** f : fs) {
		FG_CUDA(cudaMalloc(f, n * 4));
		FG_CUDA(cudaMemsetAsync(*f, 0, n * 4, h->st));
	}
	FG_CUDA(cudaMalloc(&h->cell_type, n * 4));
	FG_CUDA(cudaMemsetAsync(h->cell_type, 0xff, n * 4, h->st));  // -1: not fluid until uploaded
	FG_CUDA(cudaMalloc(&h->d_rest, 4));
	FG_CUDA(cudaMemsetAsync(h->d_rest, 0, 4, h->st));
	FG_CUDA(cudaStreamSynchronize(h->st));
	if (nranks > 1) {
		ncclUniqueId id;
		memcpy(&id, unique_id128, sizeof(id));
		FG_TRY(load_nccl());
		FG_NCCL(g_nccl.CommInitRank(&h->comm, nranks, id, rank));
	}
	*out = h;
	return FLIPGPU_OK;
}

int sorshard_destroy(SorShard* h) {
	if (!h) return FLIPGPU_OK;
	SDeviceGuard guard(h->device);
	cudaStreamSynchronize(h->st);
	if (h->comm) g_nccl.CommDestroy(h->comm);
	void* ptrs[] = {h->vel_f, h->vel_s, h->p, h->s, h->dens, h->cell_type, h->d_rest, h->alt_f, h->alt_s, h->alt_p};
	for (void* p : ptrs) if (p) cudaFree(p);
	sor_workspace_free(h->ws);
	cudaStreamDestroy(h->st);
	delete h;
	return FLIPGPU_OK;
}

int sorshard_get_rows(SorShard* h, int* j0, int* j1, int* halo) {
	FG_CHECK(h && j0 && j1 && halo, FLIPGPU_EINVAL, "null argument");
	*j0 = h->j0; *j1 = h->j1; *halo = h->halo;
	return FLIPGPU_OK;
}

// rows [j_start, j_start+n_rows) of a field, global row numbering; must lie inside own rows +- the halo rows
static int shard_copy(SorShard* h, int field, void* host, int j_start, int n_rows, bool to_device) {
	FG_CHECK(h && host, FLIPGPU_EINVAL, "null argument");
	size_t elem;
	char* base = (char*)shard_field(h, field, &elem);
	FG_CHECK(base, FLIPGPU_EINVAL, "sorshard: unknown field %d", field);
	FG_CHECK(n_rows >= 0 && j_start >= h->j0 - h->halo && j_start + n_rows <= h->j1 + h->halo && j_start >= 0 && j_start + n_rows <= h->ns,
	         FLIPGPU_EINVAL, "sorshard: rows [%d,%d) outside this rank's window [%d,%d)", j_start, j_start + n_rows, h->j0 - h->halo, h->j1 + h->halo);
	SDeviceGuard guard(h->device);
	size_t off = (size_t)(j_start - (h->j0 - h->halo)) * h->nf * elem, bytes = (size_t)n_rows * h->nf * elem;
	if (bytes) {
		if (to_device && field == SORSHARD_S) {
			const float* sv = (const float*)host;
			for (size_t i = 0, n = (size_t)n_rows * h->nf; i < n && h->s_binary; i++) h->s_binary = sv[i] == 0.f || sv[i] == 1.f;
		}
		if (to_device) FG_CUDA(cudaMemcpyAsync(base + off, host, bytes, cudaMemcpyHostToDevice, h->st));
		else FG_CUDA(cudaMemcpyAsync(host, base + off, bytes, cudaMemcpyDeviceToHost, h->st));
	}
	FG_CUDA(cudaStreamSynchronize(h->st));
	return FLIPGPU_OK;
}
int sorshard_upload_rows(SorShard* h, int field, const void* host, int j_start, int n_rows) {
	return shard_copy(h, field, const_cast<void*>(host), j_start, n_rows, true);
}
int sorshard_readback_rows(SorShard* h, int field, void* host, int j_start, int n_rows) {
	return shard_copy(h, field, host, j_start, n_rows, false);
}

// iters red-black sweeps (flip_fluid.h:458-494 update, red-black order); p is NOT cleared here
int sorshard_solve(SorShard* h, int iters, float cp, float omega, int drift, float rest_density) {
	FG_CHECK(h && iters >= 0, FLIPGPU_EINVAL, "sorshard_solve: bad argument");
	SDeviceGuard guard(h->device);
	FG_CUDA(cudaMemcpyAsync(h->d_rest, &rest_density, 4, cudaMemcpyHostToDevice, h->st));
	SorGrid g;
	g.nf = h->nf; g.ns = h->rows; g.vel_f = h->vel_f; g.vel_s = h->vel_s; g.p = h->p; g.s = h->s; g.cell_type = h->cell_type;
	g.density = h->dens; g.rest = h->d_rest; g.x_fast = h->x_fast;
	// local row l holds global row j0-H+l; interior global rows are 1..ns-2
	const int H = h->halo, K = std::max(H - 1, 1), base = h->j0 - H;
	const int shift = ((base % 2) + 2) % 2;
	const bool use_drift = drift != 0;
	const bool has_up = h->rank + 1 < h->nranks, has_down = h->rank > 0;
	dim3 block(64, 4);
	const unsigned gx = cdiv((size_t)(h->nf + 1) / 2 + 1, block.x);
	const size_t blk = (size_t)H * h->nf;  // one exchange block: H rows
	// refresh the halos: my H outermost own rows go to the neighbour's halo, its rows come into mine
	auto exchange = [&]() -> int {
		float* fields[2] = {h->vel_f, h->vel_s};
		FG_NCCL(g_nccl.GroupStart());
		for (float* f : fields) {
			if (has_up) {
				FG_NCCL(g_nccl.Send(f + (size_t)(h->rows - 2 * H) * h->nf, blk, ncclFloat, h->rank + 1, h->comm, h->st));
				FG_NCCL(g_nccl.Recv(f + (size_t)(h->rows - H) * h->nf, blk, ncclFloat, h->rank + 1, h->comm, h->st));
			}
			if (has_down) {
				FG_NCCL(g_nccl.Send(f + (size_t)H * h->nf, blk, ncclFloat, h->rank - 1, h->comm, h->st));
				FG_NCCL(g_nccl.Recv(f, blk, ncclFloat, h->rank - 1, h->comm, h->st));
			}
		}
		FG_NCCL(g_nccl.GroupEnd());
		h->exchanges++;
		return FLIPGPU_OK;
	};
	// With 0/1-valued s: ONE launch of the temporally blocked red-black kernel per exchange -- its 8 colour passes are exactly
	// what the 9-row halo allows -- as the sharded FLIP step does (flip.cu shard_solve; bit-identical to one launch per
	// colour pass, validated on 2 GPUs: tests/test_gpu_multi.py).  Rows of the local window that are the global border, or
	// outside the grid, are deactivated; output = own rows +- 1 (completes the face rows two slabs share).
	if (h->nranks > 1 && h->s_binary && K == sor_rb_passes_per_round()) {
		const size_t n = (size_t)h->nf * h->rows;
		if (!h->alt_f) {
			float** alts[] = {&h->alt_f, &h->alt_s, &h->alt_p};
			for (float** a : alts) {
				FG_CUDA(cudaMalloc(a, n * 4));
				FG_CUDA(cudaMemsetAsync(*a, 0, n * 4, h->st));
			}
		}
		FG_TRY(sor_rb_prep(&h->ws, g, use_drift, 0, h->rows, h->st, &h->lc));
		FG_TRY(sor_rb_deactivate_rows(h->ws, g, 0, 1 - base, h->st));                // global rows <= 0
		FG_TRY(sor_rb_deactivate_rows(h->ws, g, h->ns - 1 - base, h->rows, h->st));  // global rows >= ns-1
		const int out_lo = (has_down ? h->j0 - 1 : h->j0) - base, out_hi = (has_up ? h->j1 + 1 : h->j1) - base;
		for (int pass0 = 0; pass0 < 2 * iters; pass0 += K) {
			FG_TRY(exchange());
			g.vel_f = h->vel_f; g.vel_s = h->vel_s; g.p = h->p;
			FG_TRY(sor_rb_round(h->ws, g, h->alt_f, h->alt_s, h->alt_p, std::min(K, 2 * iters - pass0), (pass0 & 1) ^ shift, out_lo, out_hi,
			                    cp, omega, use_drift, h->st, &h->lc));
			std::swap(h->vel_f, h->alt_f); std::swap(h->vel_s, h->alt_s); std::swap(h->p, h->alt_p);
		}
		FG_CUDA(cudaGetLastError());
		return FLIPGPU_OK;
	}
	int pass = 0;
	for (int it = 0; it < iters; it++)
		for (int colour = 0; colour < 2; colour++, pass++) {
			const int m = h->nranks > 1 ? pass % K : 0;
			if (h->nranks > 1 && m == 0) FG_TRY(exchange());
			// rows whose inputs are still exact: own rows plus a halo band that shrinks by one row per pass
			int lo = has_down ? h->j0 - (H - 1) + m : h->j0, hi = has_up ? h->j1 + (H - 1) - m - 1 : h->j1 - 1;
			lo = std::max(lo, 1); hi = std::min(hi, h->ns - 2);
			if (hi < lo) continue;
			dim3 grid(gx, cdiv((size_t)(hi - lo + 1), block.y));
			if (h->x_fast) sor_rb_rows_kernel<true><<<grid, block, 0, h->st>>>(g, colour, shift, lo - base, hi - base, cp, omega, use_drift);
			else sor_rb_rows_kernel<false><<<grid, block, 0, h->st>>>(g, colour, shift, lo - base, hi - base, cp, omega, use_drift);
			h->lc.n++;
		}
	FG_CUDA(cudaGetLastError());
	return FLIPGPU_OK;
}

int sorshard_synchronize(SorShard* h) {
	FG_CHECK(h, FLIPGPU_EINVAL, "null handle");
	SDeviceGuard guard(h->device);
	FG_CUDA(cudaStreamSynchronize(h->st));
	return FLIPGPU_OK;
}
int sorshard_get_stream(SorShard* h, void** stream) {
	FG_CHECK(h && stream, FLIPGPU_EINVAL, "null argument");
	*stream = (void*)h->st;
	return FLIPGPU_OK;
}
int sorshard_get_counters(SorShard* h, long long* launches, long long* exchanges) {
	FG_CHECK(h && launches && exchanges, FLIPGPU_EINVAL, "null argument");
	*launches = h->lc.n; *exchanges = h->exchanges;
	return FLIPGPU_OK;
}

}  // extern "C"
