// Register-systolic wavefront Gauss-Seidel SOR: the reference's lexicographic sweep order (flip_fluid.h:458-494 /
// fluid.h:106-138), bit for bit, with the K iterations of a chunk held in REGISTERS and no block-wide barrier anywhere.
//
// Why: the shared-memory tile kernel (sor_tiled.cu) is bound by the critical path of its wavefront -- 2T-1 block-wide
// barriers per tile, 12 shared-memory accesses per cell update, ~2.3 warp instructions per update.  Here one WARP owns a
// strip of 32 grid columns (fast axis) for a chunk of K iterations and marches along the slow axis one row per step:
//
//   * lane = column mod 32; pipeline stage k = iteration k of the chunk; the strip of stage k is shifted one column to
//     the low side per stage (column F = 32w - k + c, c = (lane + k) & 31), so every "previous iteration" dependency
//     (F+1,S,k-1), (F,S+1,k-1) lies in the same warp and the only inter-warp dependency is on the strip to the left.
//   * lane (c, k) updates row S at step t = S + c + k.  Its five operands: the low row-face is the high row-face it wrote
//     one step earlier (register); the low column-face comes from lane-1 of the same stage (one rotate shuffle); the high
//     column-face from lane+1 of stage k-1 (one rotate shuffle); the high row-face, p and the cell constants are what its
//     OWN stage k-1 produced one / two steps earlier (registers).  Shared memory is touched only to skew the coalesced
//     global rows into the lanes' staggered time (stage 0 in, stage K-1 out) and at the one lane per stage where the
//     strip wraps around the warp and the hand-over crosses to the neighbour strip (rings in global memory).
//   * a step holds K independent dependency chains per lane and is straight-line code (a fast body when every active cell
//     of the step has four open neighbours, a general body otherwise), so ONE warp per scheduler keeps the issue slot busy.
//   * strips are chained by row-progress words (release/acquire), not barriers: strip w trails strip w-1 by ~32+3G rows,
//     chunk q trails chunk q-1 by about the same; the critical path is ~ (columns*1.4 + rows) steps of ~K*30 instructions.
//
// tests/systolic_model.cpp is the same dataflow on the host (lanes as array elements), checked against the plain loops.
// Compiled with -fmad=false; the only fused operations are the explicit ones of the exactly rounded x/3.
#include <algorithm>
#include <cstdlib>
#include <vector>
#include "common.cuh"

namespace fg {

constexpr int kSysG = 4;                 // steps per phase (flush / publish / dependency check / refill)
constexpr int kSysWarps = 4;             // warps per CTA, one CTA per SM: one warp per scheduler
constexpr int kSysRin = 32 + kSysG;      // rows of the skew rings (stage 0 in, stage K-1 out): window [t0-31, t0+G]
constexpr int kSysRio = 16;              // rows of the strip-to-strip rings in shared memory (needs 2G+K-2 <= 16)
constexpr int kSysPubEvery = 2;          // progress is published every this many phases
constexpr unsigned kFull = 0xffffffffu;
#ifndef SYS_STEP_UNROLL
#define SYS_STEP_UNROLL 1   // steps of a phase unrolled in the instruction stream (the two step bodies of K stages are large)
#endif
constexpr int kSysStepUnroll = SYS_STEP_UNROLL;

// cell word: bit0 active | bits1-4 s != 0 at fast-, fast+, slow-, slow+ | bit5 s_sum==3 | bit6 s_sum==2 | bit7 s_sum==1 |
// bit8 "general": active with a closed neighbour (the fast step body handles words 0 and 31 only)
__global__ void __launch_bounds__(256) sys_prep_kernel(SorGrid g, bool drift, float2* __restrict__ cd) {
	float rest = (drift && g.rest) ? *g.rest : 0.f;
	const size_t n = (size_t)g.nf * g.ns;
	for (size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x; c < n; c += (size_t)gridDim.x * blockDim.x) {
		int a = (int)(c % g.nf), b = (int)(c / g.nf);
		unsigned w = 0;
		float dr = 0.f;
		if (a >= 1 && a <= g.nf - 2 && b >= 1 && b <= g.ns - 2 && g.cell_type[c] == FLUID_CELL) {
			unsigned m = (g.s[c - 1] != 0.f ? 2u : 0u) | (g.s[c + 1] != 0.f ? 4u : 0u) | (g.s[c - g.nf] != 0.f ? 8u : 0u) |
			             (g.s[c + g.nf] != 0.f ? 16u : 0u);
			if (m) {
				int open = __popc(m);
				w = m | 1u;
				if (open == 3) w |= 32u; else if (open == 2) w |= 64u; else if (open == 1) w |= 128u;
				if (open != 4) w |= 256u;
				if (drift && rest > 0.f) {
					float compression = g.density[c] - rest;
					if (compression > 0.f) dr = compression;
				}
			}
		}
		cd[c] = make_float2(__uint_as_float(w), dr);
	}
}

struct SysArgs {
	int nf, ns, nw, nq;
	float *vf, *vs, *p;
	const float2* cd;
	const int2* tasks;   // (w, q) in start order
	int n_tasks;
	int* counter;
	int* prog;           // [nq][nw] rows complete
	int* status;
	volatile int* host_status;   // pinned host mirror of the error word, written by the kernel on a time-out
	float* ring;         // [nw][2][ns][RSF]
	float cp, omega;
	unsigned long long* prof;  // optional (FLIPGPU_SOR_PROFILE): cycles in [0] dependency waits [1] flush+publish [2] land+fetch [3] steps, [4] tasks, [5] task cycles
};

template <int K>
struct SysLayout {
	static constexpr int RSF = ((4 * (K - 1) + K) + 3) / 4 * 4;   // floats per ring row: (K-1) float4 hand-overs, then K column faces
	static constexpr int kVfOff = 4 * (K - 1);
	static constexpr int kWarpFloats = kSysRin * 32 * 5 + kSysRin * 32 * 3 + 2 * kSysRio * RSF + 32 * 4;   // skew rings in, out, strip rings, scratch
	static_assert(2 * kSysG + K - 2 <= kSysRio, "strip-to-strip ring too small");
};

template <int K, bool DRIFT>
struct SysState {
	float vs_lo[K], vf_lo[K];
	float A_p[K], A_vs[K], A_dr[K], A_vfl[K], B_p[K], B_dr[K];
	unsigned A_cd[K], B_cd[K];
};

// exactly rounded x/3 (checked against x/3.0f for all 2^32 inputs): q = RN(x*RN(1/3)) corrected by the exact residual
__device__ __forceinline__ float div3_rn(float x) {
	const float c3 = 0.3333333432674407958984375f;
	float q0 = __fmul_rn(x, c3);
	float r = __fmaf_rn(-3.0f, q0, x);
	float q1 = __fmaf_rn(r, c3, q0);
	return r == 0.f ? q0 : q1;   // keeps the sign of a zero quotient
}

// this lane's roles at the wrap of the strip and the ring addresses they use in the current step
struct SysXchg {             // (shared-memory addresses, 32-bit)
	int k31, k0;             // stage for which the lane is the last / the first column of the strip (>= K: none)
	unsigned out_vf;         // io_out row (t-31-k31): column face of stage k31 ...
	unsigned out_hand;       // ... and its float4 hand-over
	unsigned in_hand;        // io_in row (t+1-k31), float4 hand-over of stage k31
	unsigned in_vf;          // io_in row (t+1-k0), column face of stage k0
	unsigned scratch;        // this lane's private float4
};

// single predicated shared-memory accesses (ptxas keeps these predicated)
__device__ __forceinline__ void sts_if(bool p, unsigned addr, float v) {
	asm volatile("{ .reg .pred q; setp.ne.s32 q, %0, 0; @q st.shared.f32 [%1], %2; }" ::"r"((int)p), "r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float lds_if(bool p, unsigned addr, float keep) {
	asm volatile("{ .reg .pred q; setp.ne.s32 q, %1, 0; @q ld.shared.f32 %0, [%2]; }" : "+f"(keep) : "r"((int)p), "r"(addr) : "memory");
	return keep;
}

// One step of a task: every lane advances its K stages by one row.
//   compute block : straight-line, branch-free (the update is computed for every lane and committed under the cell's
//                   "active" bit), so the K dependency chains of a lane interleave in ONE basic block;
//   rotate block  : the K high column-faces go to lane+1;
//   exchange block: the (at most one) stage for which this lane is the last / first column of the strip swaps its
//                   hand-over with the strip-to-strip rings in shared memory -- short divergent branches, kept out of the
//                   compute block on purpose.
template <int K, bool XFAST, bool DRIFT, bool GEN>
__device__ __forceinline__ void sys_compute(SysState<K, DRIFT>& s, float (&vfh_out)[K], const int lane, const int lane_hi,
                                            const float* __restrict__ in_vf, const float* __restrict__ in_vs, const float* __restrict__ in_p,
                                            const float2* __restrict__ in_cd, float* __restrict__ o_p, float* __restrict__ o_vs,
                                            float* __restrict__ o_vf, const int s0, const int s1, const int sf, const float cp, const float omega) {
#pragma unroll
	for (int kk = 0; kk < K; kk++) {
		const int k = K - 1 - kk;                       // high stages first: stage k consumes what stage k-1 left LAST step
		float vfl = s.vf_lo[k], vsl = s.vs_lo[k];
		float vfh, vsh, pp, dr = 0.f;
		unsigned cd;
		if (k == 0) {
			vfh = in_vf[s0 * 32 + lane]; vsh = in_vs[s1 * 32 + lane]; pp = in_p[s0 * 32 + lane];
			float2 c2 = in_cd[s0 * 32 + lane];
			cd = __float_as_uint(c2.x);
			if (DRIFT) dr = c2.y;
		} else {
			vfh = __shfl_sync(kFull, s.A_vfl[k], lane_hi);
			vsh = s.A_vs[k]; pp = s.B_p[k]; cd = s.B_cd[k];
			if (DRIFT) dr = s.B_dr[k];
		}
		{
			const bool act = (cd & 1u) != 0u;
			float div = XFAST ? ((vfh - vfl) + vsh) - vsl : ((vsh - vsl) + vfh) - vfl;
			if (DRIFT) div -= dr;                       // the stored term is max(0, density-rest): subtracting 0 is exact
			if (!GEN) {                                 // word 31: four open neighbours, s_sum = 4, x/4 == x*0.25 exactly
				const float pv = (-div * 0.25f) * omega;
				pp = act ? pp + cp * pv : pp;
				vfl = act ? vfl - pv : vfl; vfh = act ? vfh + pv : vfh; vsl = act ? vsl - pv : vsl; vsh = act ? vsh + pv : vsh;
			} else {
				const float nd = -div;
				const float inv = (cd & 64u) ? 0.5f : ((cd & 128u) ? 1.f : 0.25f);
				const float q3 = div3_rn(nd), qi = nd * inv;
				float pv = (cd & 32u) ? q3 : qi;        // -div / s_sum
				pv *= omega;
				const float sfm = (cd & 2u) ? 1.f : 0.f, sfp = (cd & 4u) ? 1.f : 0.f, ssm = (cd & 8u) ? 1.f : 0.f, ssp = (cd & 16u) ? 1.f : 0.f;
				pp = act ? pp + cp * pv : pp;
				vfl = act ? vfl - sfm * pv : vfl; vfh = act ? vfh + sfp * pv : vfh; vsl = act ? vsl - ssm * pv : vsl; vsh = act ? vsh + ssp * pv : vsh;
			}
		}
		s.vs_lo[k] = vsh;                               // the high row-face is the low row-face of the next row
		vfh_out[k] = vfh;
		if (k >= 1) {
			s.B_p[k] = s.A_p[k]; s.B_cd[k] = s.A_cd[k];
			if (DRIFT) s.B_dr[k] = s.A_dr[k];
		}
		if (k < K - 1) {
			s.A_p[k + 1] = pp; s.A_vs[k + 1] = vsl; s.A_cd[k + 1] = cd; s.A_vfl[k + 1] = vfl;
			if (DRIFT) s.A_dr[k + 1] = dr;
		} else {
			o_p[sf * 32 + lane] = pp; o_vs[sf * 32 + lane] = vsl; o_vf[sf * 32 + lane] = vfl;
		}
	}
}

// rotate + exchange blocks, common to both compute variants (so the registers the 128-bit loads fill are the ones the next
// step reads: no copies behind the loads)
template <int K, bool DRIFT>
__device__ __forceinline__ void sys_exchange(SysState<K, DRIFT>& s, const float (&vfh_out)[K], const int lane_lo, const SysXchg& x) {
#pragma unroll
	for (int k = 0; k < K; k++) s.vf_lo[k] = __shfl_sync(kFull, vfh_out[k], lane_lo);   // the high column-face is the low column-face of lane+1
	// exchange block, branch-free.  A lane is the LAST column of at most one stage (k31 = 31 - lane) and the FIRST column of
	// at most one (k0 = (32 - lane) & 31); the ring addresses of those roles were formed by all lanes at once (x.*).  The
	// float4 hand-over is stored and re-loaded by EVERY lane: the wrap lane stores to the outgoing ring and loads from the
	// incoming one, every other lane uses a private scratch slot and reads back what it just wrote (same thread, same
	// address, program order) -- two unpredicated 128-bit accesses and two address selects per stage instead of a
	// divergent branch (ptxas turns predicated 128-bit shared accesses into branches).
#pragma unroll
	for (int k = 0; k < K; k++) {
		const bool last = x.k31 == k, first = x.k0 == k;
		sts_if(last, x.out_vf, vfh_out[k]);
		s.vf_lo[k] = lds_if(first, x.in_vf, s.vf_lo[k]);    // first column: the low column-face comes from the left strip
		if (k < K - 1) {                                    // the last column of stage k is the first column of stage k+1
			const unsigned po = last ? x.out_hand : x.scratch, pi = last ? x.in_hand : x.scratch;
			float hx = s.A_p[k + 1], hy = s.A_vs[k + 1], hz = DRIFT ? s.A_dr[k + 1] : 0.f, hw = __uint_as_float(s.A_cd[k + 1]);
			asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(po), "f"(hx), "f"(hy), "f"(hz), "f"(hw) : "memory");
			asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(hx), "=f"(hy), "=f"(hz), "=f"(hw) : "r"(pi) : "memory");
			s.A_p[k + 1] = hx; s.A_vs[k + 1] = hy; s.A_cd[k + 1] = __float_as_uint(hw);
			if (DRIFT) s.A_dr[k + 1] = hz;
		}
	}
}

// Lane 0 spins (bounded) until *pr >= need and returns the value it saw (>= need), or -1 after a time-out (which raises the
// error word).  No fence on this side: every datum the producer published is read with ld.cg (L2) AFTER the poll -- the
// loads are issued behind the branch that consumes the polled value -- and the producer's release orders its stores before
// the progress word.
__device__ __forceinline__ int sys_wait(const int* pr, int need, int* status, volatile int* host_status, int lane) {
	int seen = 0;
	if (lane == 0) {
		long long spins = 0;
		while ((seen = *((volatile const int*)pr)) < need) {
			if (++spins > (1ll << 24) || ((spins & 255) == 0 && *((volatile const int*)status) != 0)) { seen = -1; break; }
			__nanosleep(20);
		}
		if (seen < 0) { atomicExch(status, FLIPGPU_ESTATE); *host_status = FLIPGPU_ESTATE; }
	}
	return __shfl_sync(kFull, seen, 0);
}

template <int K, bool XFAST, bool DRIFT>
__global__ void __launch_bounds__(32 * kSysWarps, 1) sor_systolic_kernel(SysArgs A) {
	using L = SysLayout<K>;
	constexpr int G = kSysG;
	extern __shared__ float4 sys_smem4[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	float* sm = reinterpret_cast<float*>(sys_smem4) + (size_t)warp * L::kWarpFloats;
	float* in_vf = sm;
	float* in_vs = in_vf + kSysRin * 32;
	float* in_p = in_vs + kSysRin * 32;
	float2* in_cd = reinterpret_cast<float2*>(in_p + kSysRin * 32);
	float* o_p = in_p + kSysRin * 32 * 3;
	float* o_vs = o_p + kSysRin * 32;
	float* o_vf = o_vs + kSysRin * 32;
	float* io_in = o_vf + kSysRin * 32;
	float* io_out = io_in + kSysRio * L::RSF;
	float* scratch = io_out + kSysRio * L::RSF;
	const int nf = A.nf, ns = A.ns;
	const int lane_hi = (lane + 1) & 31, lane_lo = (lane + 31) & 31;
	const int cK = (lane + K - 1) & 31;            // this lane's position in the strip of the last stage
	const int T = ns + 31 + K;
	constexpr int kIo4 = G * L::RSF / 4;            // float4 per phase of the strip-to-strip ring
	constexpr int kIoRegs = (kIo4 + 31) / 32;

	for (;;) {
		int ti = 0;
		if (lane == 0) ti = atomicAdd(A.counter, 1);
		ti = __shfl_sync(kFull, ti, 0);
		if (ti >= A.n_tasks) return;
		if (*((volatile const int*)A.status) != 0) return;
		const int2 wq = A.tasks[ti];
		const int w = wq.x, q = wq.y;
		const int* dep_left = w > 0 ? A.prog + q * A.nw + w - 1 : nullptr;
		const int* dep_prev = q > 0 ? A.prog + (q - 1) * A.nw + min(w + 1, A.nw - 1) : nullptr;
		int* my_prog = A.prog + q * A.nw + w;
		const float* ring_in = w > 0 ? A.ring + ((size_t)(w - 1) * 2 + (q & 1)) * (size_t)ns * L::RSF : nullptr;
		float* ring_out = w + 1 < A.nw ? A.ring + ((size_t)w * 2 + (q & 1)) * (size_t)ns * L::RSF : nullptr;
		const int F0 = 32 * w + lane;                       // stage-0 column of this lane
		const int FK = 32 * w - (K - 1) + cK;               // last-stage column of this lane

		// ---- start of a task: everything inert
		for (int i = lane; i < L::kWarpFloats; i += 32) sm[i] = 0.f;
		SysState<K, DRIFT> s;
#pragma unroll
		for (int k = 0; k < K; k++) {
			s.vs_lo[k] = 0.f; s.vf_lo[k] = 0.f; s.A_p[k] = 0.f; s.A_vs[k] = 0.f; s.A_dr[k] = 0.f; s.A_vfl[k] = 0.f; s.B_p[k] = 0.f; s.B_dr[k] = 0.f;
			s.A_cd[k] = 0u; s.B_cd[k] = 0u;
		}
		__syncwarp();

		// prefetch registers: G rows of the stage-0 operands of this lane's column, G rows of the left strip's ring
		float r_vf[G], r_vs[G], r_p[G];
		float2 r_cd[G];
		float4 r_io[kIoRegs];
		auto fetch = [&](int row0) {          // rows [row0, row0+G) -> registers (L2-coherent loads: the producers are other SMs)
#pragma unroll
			for (int g = 0; g < G; g++) {
				const int r = row0 + g;
				r_vf[g] = 0.f; r_vs[g] = 0.f; r_p[g] = 0.f; r_cd[g] = make_float2(0.f, 0.f);
				if (r < ns && F0 < nf) {
					const size_t c = (size_t)F0 + (size_t)nf * r;
					if (F0 + 1 < nf) r_vf[g] = __ldcg(A.vf + c + 1);
					r_vs[g] = __ldcg(A.vs + c); r_p[g] = __ldcg(A.p + c); r_cd[g] = __ldg(A.cd + c);
				}
			}
#pragma unroll
			for (int j = 0; j < kIoRegs; j++) {
				const int i = lane + 32 * j, r = row0 + i / (L::RSF / 4);
				r_io[j] = make_float4(0.f, 0.f, 0.f, 0.f);
				if (ring_in && i < kIo4 && r < ns) r_io[j] = __ldcg(reinterpret_cast<const float4*>(ring_in) + (size_t)r * (L::RSF / 4) + i % (L::RSF / 4));
			}
		};
		auto land = [&](int row0) {           // registers -> shared memory rows [row0, row0+G)
			int slot = row0 % kSysRin;
#pragma unroll
			for (int g = 0; g < G; g++) {
				in_vf[slot * 32 + lane] = r_vf[g]; in_vs[slot * 32 + lane] = r_vs[g]; in_p[slot * 32 + lane] = r_p[g]; in_cd[slot * 32 + lane] = r_cd[g];
				slot = slot + 1 == kSysRin ? 0 : slot + 1;
			}
#pragma unroll
			for (int j = 0; j < kIoRegs; j++) {
				const int i = lane + 32 * j, r = row0 + i / (L::RSF / 4);
				if (i < kIo4) reinterpret_cast<float4*>(io_in)[(r & (kSysRio - 1)) * (L::RSF / 4) + i % (L::RSF / 4)] = r_io[j];
			}
		};
		bool alive = true;
		int have_left = dep_left ? 0 : 0x7fffffff, have_prev = dep_prev ? 0 : 0x7fffffff;   // progress of the two producers as last seen
		long long c_wait = 0, c_flush = 0, c_io = 0, c_step = 0;
		const long long c_begin = clock64();
		auto wait_rows = [&](int need) {
			need = min(need, ns);
			if (have_left >= need && have_prev >= need) return;
			const long long c0 = clock64();
			if (have_left < need) { have_left = sys_wait(dep_left, need, A.status, A.host_status, lane); alive = have_left >= 0; }
			if (alive && have_prev < need) { have_prev = sys_wait(dep_prev, need, A.status, A.host_status, lane); alive = have_prev >= 0; }
			c_wait += clock64() - c0;
		};

		// ---- prologue: row 0 directly, rows [1, G+1) into the prefetch registers
		wait_rows(G + 1);
		if (!alive) return;
		{
			float v0 = 0.f, v1 = 0.f, v2 = 0.f;
			float2 v3 = make_float2(0.f, 0.f);
			if (F0 < nf) {
				const size_t c = (size_t)F0;
				if (F0 + 1 < nf) v0 = __ldcg(A.vf + c + 1);
				v1 = __ldcg(A.vs + c); v2 = __ldcg(A.p + c); v3 = __ldg(A.cd + c);
			}
			in_vf[lane] = v0; in_vs[lane] = v1; in_p[lane] = v2; in_cd[lane] = v3;
			if (ring_in)
				for (int i = lane; i < L::RSF / 4; i += 32) reinterpret_cast<float4*>(io_in)[i] = __ldcg(reinterpret_cast<const float4*>(ring_in) + i);
		}
		fetch(1);

		// running slots of this lane: stage-0 rows S0 = t - lane and S0 + 1, last-stage row SK = t - cK - (K-1)
		SysXchg x;
		x.k31 = (31 - lane) & 31; x.k0 = (32 - lane) & 31;
		const unsigned sa_in = (unsigned)__cvta_generic_to_shared(io_in), sa_out = (unsigned)__cvta_generic_to_shared(io_out);
		x.scratch = (unsigned)__cvta_generic_to_shared(scratch + 4 * lane);
		x.out_vf = sa_out; x.out_hand = sa_out; x.in_hand = sa_in; x.in_vf = sa_in;
		int s0 = ((0 - lane) % kSysRin + kSysRin) % kSysRin;
		int s1 = s0 + 1 == kSysRin ? 0 : s0 + 1;
		int sf = ((0 - cK - (K - 1)) % kSysRin + 2 * kSysRin) % kSysRin;
		int flushed = 0, published = 0, pub_tick = 0;       // rows [0, flushed) are stored, rows [0, published) announced
		auto publish = [&](int rows) {        // release: orders the warp's earlier stores (joined by __syncwarp) before the progress word
			if (lane == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(my_prog), "r"(rows) : "memory");
			published = rows;
		};

		auto flush = [&](int done) {          // rows [flushed, done): last-stage results -> grid arrays, ring rows -> global ring
			// the rows stored at least one phase ago are announced every kSysPubEvery phases (a gpu-scope release costs an L2 round trip)
			if (flushed > published && (++pub_tick >= kSysPubEvery || done >= ns)) { publish(flushed); pub_tick = 0; }
			if (done <= flushed) return;
			__syncwarp();
			const bool col_ok = FK >= 1 && FK <= nf - 1, col_in = FK <= nf - 2;
			for (int base = flushed; base < done; base += G) {      // G rows per pass: all shared loads, then all global stores
				float fp[G], fv[G], fs[G];
				int slot = base % kSysRin;
#pragma unroll
				for (int g = 0; g < G; g++) {
					fp[g] = o_p[slot * 32 + lane]; fv[g] = o_vf[slot * 32 + lane]; fs[g] = o_vs[slot * 32 + lane];
					slot = slot + 1 == kSysRin ? 0 : slot + 1;
				}
				size_t c = (size_t)FK + (size_t)nf * base;
#pragma unroll
				for (int g = 0; g < G; g++, c += nf) {
					const int S = base + g;
					if (S >= done || S < 1 || !col_ok) continue;
					if (col_in && S <= ns - 2) __stcg(A.p + c, fp[g]);
					if (S <= ns - 2) __stcg(A.vf + c, fv[g]);
					if (col_in) __stcg(A.vs + c, fs[g]);
				}
			}
			if (ring_out) {
				const int n4 = (done - flushed) * (L::RSF / 4);
				for (int i = lane; i < n4; i += 32) {
					const int r = flushed + i / (L::RSF / 4), j = i % (L::RSF / 4);
					__stcg(reinterpret_cast<float4*>(ring_out) + (size_t)r * (L::RSF / 4) + j,
					       reinterpret_cast<const float4*>(io_out)[(r & (kSysRio - 1)) * (L::RSF / 4) + j]);
				}
			}
			flushed = done;
			__syncwarp();                     // every lane has issued its stores before lane 0 releases them (next phase)
		};

		for (int t0 = 0; t0 < T; t0 += G) {
			// ---- phase: publish what is complete, make sure the next-but-one block of rows exists, refill
			const long long c0 = clock64();
			flush(min(max(t0 - 30 - K, 0), ns));
			const long long c1 = clock64();
			wait_rows(t0 + 2 * G + 1);
			if (!alive) return;
			const long long c2 = clock64();
			land(t0 + 1);                     // rows [t0+1, t0+G+1), fetched one phase ago
			fetch(t0 + G + 1);                // rows [t0+G+1, t0+2G+1) stay in flight during the G steps below
			__syncwarp();
			const long long c3 = clock64();
			c_flush += c1 - c0; c_io += c3 - c2;
#pragma unroll kSysStepUnroll
			for (int g = 0; g < G; g++) {
				const int t = t0 + g;
				// fast body unless some active cell of this step has a closed neighbour
				unsigned any = __float_as_uint(in_cd[s0 * 32 + lane].x);
#pragma unroll
				for (int k = 1; k < K; k++) any |= s.B_cd[k];
				x.out_hand = sa_out + 4u * (unsigned)(((t - 31 - x.k31) & (kSysRio - 1)) * L::RSF + 4 * x.k31);
				x.out_vf = x.out_hand + 4u * (unsigned)(L::kVfOff - 3 * x.k31);
				x.in_hand = sa_in + 4u * (unsigned)(((t + 1 - x.k31) & (kSysRio - 1)) * L::RSF + 4 * x.k31);
				x.in_vf = sa_in + 4u * (unsigned)(((t + 1 - x.k0) & (kSysRio - 1)) * L::RSF + L::kVfOff + x.k0);
				float vfh_out[K];
				if (__any_sync(kFull, any & 256u))
					sys_compute<K, XFAST, DRIFT, true>(s, vfh_out, lane, lane_hi, in_vf, in_vs, in_p, in_cd, o_p, o_vs, o_vf, s0, s1, sf, A.cp, A.omega);
				else
					sys_compute<K, XFAST, DRIFT, false>(s, vfh_out, lane, lane_hi, in_vf, in_vs, in_p, in_cd, o_p, o_vs, o_vf, s0, s1, sf, A.cp, A.omega);
				sys_exchange<K, DRIFT>(s, vfh_out, lane_lo, x);
				s0 = s1;
				s1 = s1 + 1 == kSysRin ? 0 : s1 + 1;
				sf = sf + 1 == kSysRin ? 0 : sf + 1;
			}
			c_step += clock64() - c3;
		}
		flush(ns);
		publish(ns);
		if (A.prof && lane == 0) {
			atomicAdd(A.prof + 0, (unsigned long long)c_wait); atomicAdd(A.prof + 1, (unsigned long long)c_flush); atomicAdd(A.prof + 2, (unsigned long long)c_io);
			atomicAdd(A.prof + 3, (unsigned long long)c_step); atomicAdd(A.prof + 4, 1ull); atomicAdd(A.prof + 5, (unsigned long long)(clock64() - c_begin));
		}
	}
}

// ------------------------------------------------------------------------------------------------ host side
struct SysWorkspace {
	int nf = 0, ns = 0;
	float2* cd = nullptr;
	float* ring = nullptr;
	size_t ring_floats = 0;
	int2* tasks = nullptr;
	int tasks_cap = 0, tasks_K = 0, tasks_nq = 0;
	int* prog = nullptr;
	int prog_cap = 0;
	int* counter = nullptr;
	int* status = nullptr;
	int* h_status = nullptr;
	unsigned long long* prof = nullptr;
	int sms = 0;
};

void sys_workspace_free(SysWorkspace* w) {
	if (!w) return;
	cudaFree(w->cd); cudaFree(w->ring); cudaFree(w->tasks); cudaFree(w->prog); cudaFree(w->counter); cudaFree(w->status); cudaFree(w->prof);
	if (w->h_status) cudaFreeHost(w->h_status);
	delete w;
}
int sys_workspace_profile(SysWorkspace* w, unsigned long long out[8]) {
	for (int i = 0; i < 8; i++) out[i] = 0;
	if (!w || !w->prof) return FLIPGPU_ESTATE;
	FG_CUDA(cudaMemcpy(out, w->prof, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
	FG_CUDA(cudaMemset(w->prof, 0, 8 * sizeof(unsigned long long)));
	return FLIPGPU_OK;
}
int sys_workspace_check(SysWorkspace* w) {
	if (!w || !w->h_status || *w->h_status == 0) return FLIPGPU_OK;
	fg::set_error("pressure solve: a strip-to-strip dependency wait timed out (GPU preempted or oversubscribed?); the solve is incomplete");
	return FLIPGPU_ESTATE;
}

template <int K>
static int sys_launch(SysWorkspace* w, const SorGrid& g, int chunks, float cp, float omega, bool use_drift, cudaStream_t st, LaunchCounter* lc) {
	using L = SysLayout<K>;
	const int nw = (g.nf + K - 1 + 31) / 32;
	const int n_tasks = nw * chunks;
	const size_t ring_floats = (size_t)nw * 2 * g.ns * L::RSF;
	if (ring_floats > w->ring_floats) {
		cudaFree(w->ring);
		w->ring = nullptr;
		FG_CUDA(cudaMalloc(&w->ring, ring_floats * 4));
		w->ring_floats = ring_floats;
	}
	if (n_tasks > w->prog_cap) {
		cudaFree(w->prog);
		w->prog = nullptr;
		FG_CUDA(cudaMalloc(&w->prog, (size_t)n_tasks * 4));
		w->prog_cap = n_tasks;
	}
	if (w->tasks_K != K || w->tasks_nq != chunks) {
		// start order: strip w of chunk q can begin ~Ls steps after strip w-1 and ~Lc steps after strip w+1 of chunk q-1
		const int Ls = 2 * kSysG + 31 + K + 4, Lc = Ls;
		std::vector<std::pair<int, int2>> keyed;
		keyed.reserve(n_tasks);
		for (int q = 0; q < chunks; q++)
			for (int ww = 0; ww < nw; ww++) keyed.push_back({Ls * ww + (Ls + Lc) * q, make_int2(ww, q)});
		std::stable_sort(keyed.begin(), keyed.end(), [](const std::pair<int, int2>& a, const std::pair<int, int2>& b) { return a.first < b.first; });
		std::vector<int2> tasks(n_tasks);
		for (int i = 0; i < n_tasks; i++) tasks[i] = keyed[i].second;
		if (n_tasks > w->tasks_cap) {
			cudaFree(w->tasks);
			w->tasks = nullptr;
			FG_CUDA(cudaMalloc(&w->tasks, (size_t)n_tasks * sizeof(int2)));
			w->tasks_cap = n_tasks;
		}
		FG_CUDA(cudaMemcpyAsync(w->tasks, tasks.data(), (size_t)n_tasks * sizeof(int2), cudaMemcpyHostToDevice, st));
		FG_CUDA(cudaStreamSynchronize(st));  // tasks is a host temporary
		w->tasks_K = K; w->tasks_nq = chunks;
	}
	FG_CUDA(cudaMemsetAsync(w->prog, 0, (size_t)n_tasks * 4, st));
	FG_CUDA(cudaMemsetAsync(w->counter, 0, 4, st));
	SysArgs A;
	A.nf = g.nf; A.ns = g.ns; A.nw = nw; A.nq = chunks; A.vf = g.vel_f; A.vs = g.vel_s; A.p = g.p; A.cd = w->cd; A.tasks = w->tasks;
	A.n_tasks = n_tasks; A.counter = w->counter; A.prog = w->prog; A.status = w->status; A.host_status = w->h_status; A.ring = w->ring; A.cp = cp; A.omega = omega; A.prof = w->prof;
	const void* fn = g.x_fast ? (use_drift ? (const void*)sor_systolic_kernel<K, true, true> : (const void*)sor_systolic_kernel<K, true, false>)
	                          : (use_drift ? (const void*)sor_systolic_kernel<K, false, true> : (const void*)sor_systolic_kernel<K, false, false>);
	const size_t smem = (size_t)kSysWarps * L::kWarpFloats * 4;
	FG_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	const int blocks = std::min((n_tasks + kSysWarps - 1) / kSysWarps, w->sms);
	void* args[] = {&A};
	FG_CUDA(cudaLaunchKernel(fn, dim3(blocks), dim3(32 * kSysWarps), args, smem, st));
	if (lc) lc->n++;
	return FLIPGPU_OK;
}

// s must be 0/1-valued (checked by the caller).  Iterations are run in chunks of 10 / 5 / 2 / 1 pipeline stages.
int sor_solve_systolic(SysWorkspace** wsp, const SorGrid& g, int iters, float cp, float omega, bool drift, cudaStream_t st, LaunchCounter* lc) {
	if (iters <= 0 || g.nf < 3 || g.ns < 3) return FLIPGPU_OK;
	if (!*wsp) *wsp = new SysWorkspace;
	SysWorkspace* w = *wsp;
	if (w->nf != g.nf || w->ns != g.ns) {
		cudaFree(w->cd);
		w->cd = nullptr;
		FG_CUDA(cudaMalloc(&w->cd, (size_t)g.nf * g.ns * sizeof(float2)));
		w->nf = g.nf; w->ns = g.ns; w->tasks_K = 0;
		if (!w->counter) {
			FG_CUDA(cudaMalloc(&w->counter, 4));
			FG_CUDA(cudaMalloc(&w->status, 4));
			FG_CUDA(cudaMemset(w->status, 0, 4));
			FG_CUDA(cudaMallocHost(&w->h_status, 4));
			*w->h_status = 0;
			w->sms = num_sms();
			if (getenv("FLIPGPU_SOR_PROFILE")) {
				FG_CUDA(cudaMalloc(&w->prof, 8 * sizeof(unsigned long long)));
				FG_CUDA(cudaMemset(w->prof, 0, 8 * sizeof(unsigned long long)));
			}
		}
	}
	const bool use_drift = drift && g.density && g.rest;
	sys_prep_kernel<<<wave_grid((size_t)g.nf * g.ns, 256), 256, 0, st>>>(g, use_drift, w->cd);
	if (lc) lc->n++;
	FG_CUDA(cudaGetLastError());
	// chunk depth: deep chunks (10) amortise the skew rings best, shallow ones (5) shorten the critical path of the strip
	// pipeline; take the shallow ones while the deep schedule would leave warps without a task
	int kmax = 10;
	const int nw10 = (g.nf + 9 + 31) / 32;
	if ((long long)nw10 * ((iters + 9) / 10) < 2ll * w->sms * kSysWarps) kmax = 5;
	if (const char* e = getenv("FLIPGPU_SYS_K")) kmax = atoi(e) >= 10 ? 10 : 5;   // tuning knob
	int left = iters;
	const int depths[4] = {10, 5, 2, 1};
	for (int d : depths) {
		if (d > kmax || left < d) continue;
		const int chunks = left / d;
		left -= chunks * d;
		switch (d) {
			case 10: FG_TRY(sys_launch<10>(w, g, chunks, cp, omega, use_drift, st, lc)); break;
			case 5: FG_TRY(sys_launch<5>(w, g, chunks, cp, omega, use_drift, st, lc)); break;
			case 2: FG_TRY(sys_launch<2>(w, g, chunks, cp, omega, use_drift, st, lc)); break;
			default: FG_TRY(sys_launch<1>(w, g, chunks, cp, omega, use_drift, st, lc)); break;
		}
	}
		return FLIPGPU_OK;
}

}  // namespace fg
