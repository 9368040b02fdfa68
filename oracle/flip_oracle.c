/* TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
 *
 * Plain-C CPU restatement ("port") of the reference FLIP step,
 *   /root/reference/flip_fluid/src/flip_fluid.h  (struct FlipFluid).
 * Each function cites the reference lines it follows.  It is pinned
 * bit-for-bit against the reference itself (oracle/_ref/libref_flip.so, the
 * unmodified header compiled by oracle/Makefile) in tests/test_oracle_pin.py
 * and against the committed fixtures in tests/golden/ (made from the
 * reference by tests/golden/make_golden.py).  The reference has no tests or
 * golden vectors of its own, so "pinned" means: pinned to outputs of
 * the reference run in the build container.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared   (no -march=native, no
 * -ffast-math: fp32 ops must stay unfused and in source order).
 * Conventions adopted where the reference is undefined:
 * all state zero-initialised; cell_type reads below index 0 see "not fluid".
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

enum { FLUID_CELL = 0, AIR_CELL = 1, SOLID_CELL = 2 }; /* flip_fluid.h:30-34 */

typedef struct {
	float density;
	int nx, ny, ncells;
	float h, inv_h;
	float *u, *v, *du, *dv, *prev_u, *prev_v, *p, *s;
	int* cell_type; /* points past a guard of nx+1 SOLID ints */
	int* cell_type_alloc;
	float* cell_color;
	int max_particles;
	float *ppos, *pcol, *pvel, *pdens;
	float rest_density;
	float radius, p_inv;
	int pnx, pny, pncells;
	int *ncp, *fcp, *ids;
	int np;
} Flip;

static int clampi(int a, int lo, int hi) { return a < lo ? lo : (a > hi ? hi : a); }
static float clampf(float a, float lo, float hi) { return a < lo ? lo : (a > hi ? hi : a); }
static int mini(int a, int b) { return a < b ? a : b; }
static int maxi(int a, int b) { return a > b ? a : b; }

/* ctor, flip_fluid.h:62-113.  All sizes in fp32 exactly as written there. */
void* orcflip_create(float density, float width, float height, float sp, float r, int maxp) {
	Flip* f = (Flip*)calloc(1, sizeof(Flip));
	f->density = density;
	f->nx = (int)(1 + width / sp);  /* :69 */
	f->ny = (int)(1 + height / sp); /* :70 */
	f->ncells = f->nx * f->ny;
	{
		float a = width / f->nx, b = height / f->ny; /* :73 */
		f->h = a > b ? a : b;
	}
	f->inv_h = 1 / f->h; /* :74 */
	size_t C = (size_t)f->ncells;
	f->u = calloc(C, 4); f->v = calloc(C, 4); f->du = calloc(C, 4); f->dv = calloc(C, 4);
	f->prev_u = calloc(C, 4); f->prev_v = calloc(C, 4); f->p = calloc(C, 4); f->s = calloc(C, 4);
	f->cell_type_alloc = calloc(C + f->nx + 1, 4);
	for (int i = 0; i < f->nx + 1; i++) f->cell_type_alloc[i] = SOLID_CELL;
	f->cell_type = f->cell_type_alloc + f->nx + 1;
	f->cell_color = calloc(3 * C, 4);
	f->max_particles = maxp;
	f->ppos = calloc(2 * (size_t)maxp, 4);
	f->pcol = calloc(3 * (size_t)maxp, 4);
	for (int i = 0; i < maxp; i++) f->pcol[2 + 3 * i] = 1; /* :94-96 */
	f->pvel = calloc(2 * (size_t)maxp, 4);
	f->pdens = calloc(C, 4);
	f->rest_density = 0;
	f->radius = r;
	f->p_inv = 1 / (2.2f * r);              /* :103 */
	f->pnx = (int)(1 + width * f->p_inv);   /* :104 */
	f->pny = (int)(1 + height * f->p_inv);  /* :105 */
	f->pncells = f->pnx * f->pny;
	f->ncp = calloc((size_t)f->pncells, 4);
	f->fcp = calloc((size_t)f->pncells + 1, 4);
	f->ids = calloc((size_t)maxp, 4);
	f->np = 0;
	return f;
}

void orcflip_destroy(void* h) {
	Flip* f = (Flip*)h;
	free(f->u); free(f->v); free(f->du); free(f->dv); free(f->prev_u); free(f->prev_v);
	free(f->p); free(f->s); free(f->cell_type_alloc); free(f->cell_color);
	free(f->ppos); free(f->pcol); free(f->pvel); free(f->pdens);
	free(f->ncp); free(f->fcp); free(f->ids);
	free(f);
}

void orcflip_dims(void* h, int* ints, float* floats) {
	Flip* f = (Flip*)h;
	ints[0] = f->nx; ints[1] = f->ny; ints[2] = f->ncells;
	ints[3] = f->pnx; ints[4] = f->pny; ints[5] = f->pncells;
	ints[6] = f->max_particles; ints[7] = f->np;
	floats[0] = f->density; floats[1] = f->h; floats[2] = f->inv_h;
	floats[3] = f->radius; floats[4] = f->p_inv; floats[5] = f->rest_density;
}

void* orcflip_ptr(void* h, const char* name) {
	Flip* f = (Flip*)h;
#define FIELD(n, m) if (!strcmp(name, n)) return (void*)f->m;
	FIELD("u", u) FIELD("v", v) FIELD("du", du) FIELD("dv", dv) FIELD("prev_u", prev_u) FIELD("prev_v", prev_v)
	FIELD("p", p) FIELD("s", s) FIELD("cell_type", cell_type) FIELD("cell_color", cell_color)
	FIELD("particle_pos", ppos) FIELD("particle_color", pcol) FIELD("particle_vel", pvel)
	FIELD("particle_density", pdens) FIELD("num_cell_particles", ncp) FIELD("first_cell_particle", fcp)
	FIELD("cell_particle_ids", ids)
#undef FIELD
	return NULL;
}

void orcflip_set_num_particles(void* h, int n) { ((Flip*)h)->np = n; }
void orcflip_set_rest_density(void* h, float d) { ((Flip*)h)->rest_density = d; }

/* integrateParticles, flip_fluid.h:156-162 */
void orcflip_integrate(void* h, float dt, float g) {
	Flip* f = (Flip*)h;
	for (int i = 0; i < f->np; i++) {
		f->pvel[2 * i + 1] += g * dt;
		f->ppos[2 * i] += f->pvel[2 * i] * dt;
		f->ppos[2 * i + 1] += f->pvel[2 * i + 1] * dt;
	}
}

/* binning half of pushParticlesApart, flip_fluid.h:169-198 (Q7: ids descend inside a cell) */
static int hash_cell(const Flip* f, float x, float y) {
	int xi = clampi((int)(x * f->p_inv), 0, f->pnx - 1);
	int yi = clampi((int)(y * f->p_inv), 0, f->pny - 1);
	return xi + f->pnx * yi;
}
void orcflip_bin(void* h) {
	Flip* f = (Flip*)h;
	memset(f->ncp, 0, sizeof(int) * (size_t)f->pncells);
	for (int i = 0; i < f->np; i++) f->ncp[hash_cell(f, f->ppos[2 * i], f->ppos[2 * i + 1])]++;
	int run = 0;
	for (int c = 0; c < f->pncells; c++) { run += f->ncp[c]; f->fcp[c] = run; }
	f->fcp[f->pncells] = run;
	for (int i = 0; i < f->np; i++) {
		int c = hash_cell(f, f->ppos[2 * i], f->ppos[2 * i + 1]);
		f->ids[--f->fcp[c]] = i;
	}
}

/* separation sweeps, flip_fluid.h:201-250.  Q1: the partner only moves in x; the local py
 * absorbs the y push.  Q6: bins are not rebuilt between sweeps. */
void orcflip_separate(void* h, int num_iter) {
	Flip* f = (Flip*)h;
	const float diffusion = .001f; /* :166 */
	float min_dist = 2 * f->radius;
	float min_dist_sq = min_dist * min_dist;
	float* pos = f->ppos;
	float* col = f->pcol;
	for (int it = 0; it < num_iter; it++) {
		for (int i = 0; i < f->np; i++) {
			float px = pos[2 * i], py = pos[2 * i + 1];
			int pxi = (int)(px * f->p_inv), pyi = (int)(py * f->p_inv);
			int x0 = maxi(pxi - 1, 0), y0 = maxi(pyi - 1, 0);
			int x1 = mini(pxi + 1, f->pnx - 1), y1 = mini(pyi + 1, f->pny - 1);
			for (int xi = x0; xi <= x1; xi++)
				for (int yi = y0; yi <= y1; yi++) {
					int c = xi + f->pnx * yi;
					for (int k = f->fcp[c]; k < f->fcp[c + 1]; k++) {
						int q = f->ids[k];
						if (q == i) continue;
						float dx = pos[2 * q] - px, dy = pos[2 * q + 1] - py;
						float d2 = dx * dx + dy * dy;
						if (d2 > min_dist_sq || d2 == 0) continue;
						float d = sqrtf(d2);
						float sc = .5f * (min_dist - d) / d;
						dx *= sc; dy *= sc;
						pos[2 * i] -= dx;
						pos[2 * i + 1] -= dy;
						pos[2 * q] += dx;
						py += dy;
						for (int ch = 0; ch < 3; ch++) {
							float c0 = col[ch + 3 * i], c1 = col[ch + 3 * q];
							float mean = (c0 + c1) / 2;
							col[ch + 3 * i] = c0 + diffusion * (mean - c0);
							col[ch + 3 * q] = c1 + diffusion * (mean - c1);
						}
					}
				}
		}
	}
}
void orcflip_push_apart(void* h, int iters) { orcflip_bin(h); orcflip_separate(h, iters); }

/* handleParticleCollisions, flip_fluid.h:254-289 (Q12: obstacle only overwrites velocity) */
void orcflip_collide(void* h, float ox, float oy, float ovx, float ovy, float orad) {
	Flip* f = (Flip*)h;
	float md = f->radius + orad, md2 = md * md;
	float min_x = f->h + f->radius, max_x = f->h * (f->nx - 1) - f->radius;
	float min_y = f->h + f->radius, max_y = f->h * (f->ny - 1) - f->radius;
	for (int i = 0; i < f->np; i++) {
		float x = f->ppos[2 * i], y = f->ppos[2 * i + 1];
		float vx = f->pvel[2 * i], vy = f->pvel[2 * i + 1];
		float dx = x - ox, dy = y - oy;
		if (dx * dx + dy * dy < md2) { vx = ovx; vy = ovy; }
		if (x < min_x) { x = min_x; vx = 0; }
		if (x > max_x) { x = max_x; vx = 0; }
		if (y < min_y) { y = min_y; vy = 0; }
		if (y > max_y) { y = max_y; vy = 0; }
		f->ppos[2 * i] = x; f->ppos[2 * i + 1] = y;
		f->pvel[2 * i] = vx; f->pvel[2 * i + 1] = vy;
	}
}

/* bilinear stencil shared by :301-313 and :374-396 (Q2: cell-centred offset for u AND v) */
typedef struct { int n[4]; float w[4]; } Stencil;
static Stencil stencil(const Flip* f, float x, float y) {
	float h = f->h, h2 = .5f * h;
	x = clampf(x, h, h * (f->nx - 1));
	y = clampf(y, h, h * (f->ny - 1));
	int x0 = (int)(f->inv_h * (x - h2));
	float tx = f->inv_h * (x - h2 - h * x0);
	int x1 = mini(1 + x0, f->nx - 2);
	int y0 = (int)(f->inv_h * (y - h2));
	float ty = f->inv_h * (y - h2 - h * y0);
	int y1 = mini(1 + y0, f->ny - 2);
	float sx = 1 - tx, sy = 1 - ty;
	Stencil st;
	st.w[0] = sx * sy; st.w[1] = tx * sy; st.w[2] = tx * ty; st.w[3] = sx * ty;
	st.n[0] = x0 + f->nx * y0; st.n[1] = x1 + f->nx * y0; st.n[2] = x1 + f->nx * y1; st.n[3] = x0 + f->nx * y1;
	return st;
}

/* updateParticleDensity, flip_fluid.h:292-333 (add order x0y0,x1y0,x0y1,x1y1; Q14 latch) */
void orcflip_density(void* h) {
	Flip* f = (Flip*)h;
	for (int c = 0; c < f->ncells; c++) f->pdens[c] = 0;
	for (int i = 0; i < f->np; i++) {
		Stencil st = stencil(f, f->ppos[2 * i], f->ppos[2 * i + 1]);
		f->pdens[st.n[0]] += st.w[0];
		f->pdens[st.n[1]] += st.w[1];
		f->pdens[st.n[3]] += st.w[3];
		f->pdens[st.n[2]] += st.w[2];
	}
	if (f->rest_density == 0) {
		float sum = 0; int cnt = 0;
		for (int c = 0; c < f->ncells; c++)
			if (f->cell_type[c] == FLUID_CELL) { sum += f->pdens[c]; cnt++; }
		if (cnt > 0) f->rest_density = sum / cnt;
	}
}

/* transferVelocities, flip_fluid.h:336-448 */
void orcflip_transfer(void* h, int to_grid, float flip_ratio) {
	Flip* f = (Flip*)h;
	int nx = f->nx, ny = f->ny, C = f->ncells;
	if (to_grid) { /* :339-360 */
		memcpy(f->prev_u, f->u, 4 * (size_t)C);
		memcpy(f->prev_v, f->v, 4 * (size_t)C);
		for (int c = 0; c < C; c++) { f->du[c] = 0; f->dv[c] = 0; f->u[c] = 0; f->v[c] = 0; }
		for (int c = 0; c < C; c++) f->cell_type[c] = f->s[c] == 0 ? SOLID_CELL : AIR_CELL;
		for (int i = 0; i < f->np; i++) {
			int xi = clampi((int)(f->inv_h * f->ppos[2 * i]), 0, nx - 1);
			int yi = clampi((int)(f->inv_h * f->ppos[2 * i + 1]), 0, ny - 1);
			if (f->cell_type[xi + nx * yi] == AIR_CELL) f->cell_type[xi + nx * yi] = FLUID_CELL;
		}
	}
	for (int comp = 0; comp < 2; comp++) {
		float* fld = comp == 0 ? f->u : f->v;
		float* prev = comp == 0 ? f->prev_u : f->prev_v;
		float* wsum = comp == 0 ? f->du : f->dv;
		int offset = comp == 0 ? 1 : nx; /* :405 */
		for (int i = 0; i < f->np; i++) {
			Stencil st = stencil(f, f->ppos[2 * i], f->ppos[2 * i + 1]);
			if (to_grid) { /* :398-403 */
				float pv = f->pvel[comp + 2 * i];
				for (int k = 0; k < 4; k++) { fld[st.n[k]] += pv * st.w[k]; wsum[st.n[k]] += st.w[k]; }
			} else { /* :404-429 */
				float valid[4];
				for (int k = 0; k < 4; k++)
					valid[k] = (f->cell_type[st.n[k]] == FLUID_CELL || f->cell_type[st.n[k] - offset] == FLUID_CELL) ? 1.f : 0.f;
				float dsum = valid[0] * st.w[0] + valid[1] * st.w[1] + valid[2] * st.w[2] + valid[3] * st.w[3];
				if (dsum > 0) {
					float pic = (valid[0] * st.w[0] * fld[st.n[0]] + valid[1] * st.w[1] * fld[st.n[1]] +
					             valid[2] * st.w[2] * fld[st.n[2]] + valid[3] * st.w[3] * fld[st.n[3]]) / dsum;
					float corr = (valid[0] * st.w[0] * (fld[st.n[0]] - prev[st.n[0]]) +
					              valid[1] * st.w[1] * (fld[st.n[1]] - prev[st.n[1]]) +
					              valid[2] * st.w[2] * (fld[st.n[2]] - prev[st.n[2]]) +
					              valid[3] * st.w[3] * (fld[st.n[3]] - prev[st.n[3]])) / dsum;
					float pvv = f->pvel[comp + 2 * i];
					float flipv = pvv + corr;
					f->pvel[comp + 2 * i] = (1 - flip_ratio) * pic + flip_ratio * flipv;
				}
			}
		}
		if (to_grid) { /* :432-446 (Q8: both restores run in each component pass) */
			for (int c = 0; c < C; c++) if (wsum[c] > 0) fld[c] /= wsum[c];
			for (int i = 0; i < nx; i++)
				for (int j = 0; j < ny; j++) {
					int c = i + nx * j;
					int solid = f->cell_type[c] == SOLID_CELL;
					if (solid || (i > 0 && f->cell_type[c - 1] == SOLID_CELL)) f->u[c] = f->prev_u[c];
					if (solid || (j > 0 && f->cell_type[c - nx] == SOLID_CELL)) f->v[c] = f->prev_v[c];
				}
		}
	}
}

/* solveIncompressibility, flip_fluid.h:451-495 (Q9: i outer, j inner over x-fastest storage) */
void orcflip_solve(void* h, int iters, float dt, float omega, int drift) {
	Flip* f = (Flip*)h;
	int nx = f->nx, ny = f->ny, C = f->ncells;
	for (int c = 0; c < C; c++) f->p[c] = 0;
	memcpy(f->prev_u, f->u, 4 * (size_t)C);
	memcpy(f->prev_v, f->v, 4 * (size_t)C);
	float cp = f->density * f->h / dt;
	for (int it = 0; it < iters; it++)
		for (int i = 1; i < nx - 1; i++)
			for (int j = 1; j < ny - 1; j++) {
				int c = i + nx * j;
				if (f->cell_type[c] != FLUID_CELL) continue;
				int l = c - 1, r = c + 1, b = c - nx, t = c + nx;
				float sx0 = f->s[l], sx1 = f->s[r], sy0 = f->s[b], sy1 = f->s[t];
				float ssum = sx0 + sx1 + sy0 + sy1;
				if (ssum == 0) continue;
				float div = f->u[r] - f->u[c] + f->v[t] - f->v[c];
				if (f->rest_density > 0 && drift) {
					float k = 1;
					float comp = f->pdens[c] - f->rest_density;
					if (comp > 0) div -= k * comp;
				}
				float pv = -div / ssum;
				pv *= omega;
				f->p[c] += cp * pv;
				f->u[c] -= sx0 * pv;
				f->u[r] += sx1 * pv;
				f->v[c] -= sy0 * pv;
				f->v[t] += sy1 * pv;
			}
}

/* updateParticleColors, flip_fluid.h:498-518 */
void orcflip_particle_colors(void* h) {
	Flip* f = (Flip*)h;
	for (int i = 0; i < f->np; i++) {
		f->pcol[3 * i] = clampf(f->pcol[3 * i] - .01f, 0.f, 1.f);
		f->pcol[3 * i + 1] = clampf(f->pcol[3 * i + 1] - .01f, 0.f, 1.f);
		f->pcol[3 * i + 2] = clampf(f->pcol[3 * i + 2] + .01f, 0.f, 1.f);
		if (f->rest_density > 0) {
			int xi = clampi((int)(f->inv_h * f->ppos[2 * i]), 1, f->nx - 1);
			int yi = clampi((int)(f->inv_h * f->ppos[2 * i + 1]), 1, f->ny - 1);
			float rel = f->pdens[xi + f->nx * yi] / f->rest_density;
			if (rel < .7f) { f->pcol[3 * i] = .8f; f->pcol[3 * i + 1] = .8f; f->pcol[3 * i + 2] = 1; }
		}
	}
}

/* setSciColor :521-540 + updateCellColors :543-557 */
void orcflip_cell_colors(void* h) {
	Flip* f = (Flip*)h;
	for (int c = 0; c < 3 * f->ncells; c++) f->cell_color[c] = 0;
	for (int c = 0; c < f->ncells; c++) {
		if (f->cell_type[c] == SOLID_CELL) {
			f->cell_color[3 * c] = .5f; f->cell_color[3 * c + 1] = .5f; f->cell_color[3 * c + 2] = .5f;
		} else if (f->cell_type[c] == FLUID_CELL) {
			float val = f->pdens[c];
			if (f->rest_density > 0) val /= f->rest_density;
			float lo = 0, hi = 2;
			val = clampf(val, lo, hi - .0001f);
			float span = hi - lo;
			val = span == 0 ? .5f : (val - lo) / span;
			float m = .25f;
			int num = (int)(val / m);
			float sc = (val - num * m) / m;
			float r = 0, g = 0, b = 0;
			switch (num) {
				case 0: r = 0; g = sc; b = 1; break;
				case 1: r = 0; g = 1; b = 1 - sc; break;
				case 2: r = sc; g = 1; b = 0; break;
				case 3: r = 1; g = 1 - sc; b = 0; break;
			}
			f->cell_color[3 * c] = r; f->cell_color[3 * c + 1] = g; f->cell_color[3 * c + 2] = b;
		}
	}
}

static double now_ns(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return 1e9 * (double)ts.tv_sec + (double)ts.tv_nsec;
}

/* simulate, flip_fluid.h:560-581 (Q11: one sub-step).  phase_ns may be NULL.
 * phase_ns: 0 integrate 1 pushApart 2 collide 3 P2G 4 density 5 solve 6 G2P 7 colours */
void orcflip_simulate_timed(void* h, float dt, float gravity, float flip_ratio, int num_pressure_iters,
                            int num_particle_iters, float over_relaxation, int compensate_drift, int separate_particles,
                            float ox, float oy, float ovx, float ovy, float orad, int n_steps, double* phase_ns) {
	double dummy[8];
	double* t = phase_ns ? phase_ns : dummy;
	for (int k = 0; k < n_steps; k++) {
		double a = now_ns();
		orcflip_integrate(h, dt, gravity);
		double b = now_ns(); t[0] += b - a; a = b;
		if (separate_particles) orcflip_push_apart(h, num_particle_iters);
		b = now_ns(); t[1] += b - a; a = b;
		orcflip_collide(h, ox, oy, ovx, ovy, orad);
		b = now_ns(); t[2] += b - a; a = b;
		orcflip_transfer(h, 1, .5f);
		b = now_ns(); t[3] += b - a; a = b;
		orcflip_density(h);
		b = now_ns(); t[4] += b - a; a = b;
		orcflip_solve(h, num_pressure_iters, dt, over_relaxation, compensate_drift);
		b = now_ns(); t[5] += b - a; a = b;
		orcflip_transfer(h, 0, flip_ratio);
		b = now_ns(); t[6] += b - a; a = b;
		orcflip_particle_colors(h);
		orcflip_cell_colors(h);
		b = now_ns(); t[7] += b - a;
	}
}
void orcflip_simulate(void* h, float dt, float gravity, float flip_ratio, int num_pressure_iters,
                      int num_particle_iters, float over_relaxation, int compensate_drift, int separate_particles,
                      float ox, float oy, float ovx, float ovy, float orad, int n_steps) {
	orcflip_simulate_timed(h, dt, gravity, flip_ratio, num_pressure_iters, num_particle_iters, over_relaxation,
	                       compensate_drift, separate_particles, ox, oy, ovx, ovy, orad, n_steps, NULL);
}

/* ---------------------------------------------------------------------------------------------
 * CPU model of the PRODUCT's parallel separation schedule (not in the reference).
 * The reference sweep (orcflip_separate above, flip_fluid.h:201-250) is a serial Gauss-Seidel with
 * an asymmetric pair update and has no parallel equivalent; the CUDA kernel k_separate runs one
 * Jacobi sweep per launch instead.  This function is that schedule restated on the CPU so the
 * kernel can be checked BIT-EXACTLY against it, and so its statistical distance to the reference
 * sweep can be measured on the CPU (tests/test_separation_model.py):
 *   for every particle i, from the positions at the start of the sweep: window from the current
 *   position (:208-213), rows yi ascending, slots of the row segment ascending (= pIX order,
 *   descending id inside a cell), symmetric half-overlap push of :231-235 applied to i only.
 * Requires the bins of orcflip_bin (built once before the first sweep, Q6). */
void orcflip_separate_jacobi(void* h, int num_iter, int with_colors) {
	Flip* f = (Flip*)h;
	float min_dist = 2 * f->radius, min_dist_sq = min_dist * min_dist;
	size_t n2 = 2 * (size_t)f->np, n3 = 3 * (size_t)f->np;
	float* next = malloc(4 * (n2 ? n2 : 1));
	float* ncol = malloc(4 * (n3 ? n3 : 1));
	for (int it = 0; it < num_iter; it++) {
		const float* pos = f->ppos;
		for (int i = 0; i < f->np; i++) {
			float px = pos[2 * i], py = pos[2 * i + 1];
			float ax = px, ay = py;
			float c0 = f->pcol[3 * i], c1 = f->pcol[3 * i + 1], c2 = f->pcol[3 * i + 2];
			int pxi = (int)(px * f->p_inv), pyi = (int)(py * f->p_inv);
			int x0 = maxi(pxi - 1, 0), y0 = maxi(pyi - 1, 0);
			int x1 = mini(pxi + 1, f->pnx - 1), y1 = mini(pyi + 1, f->pny - 1);
			if (x0 <= x1)
				for (int yi = y0; yi <= y1; yi++)
					for (int k = f->fcp[x0 + f->pnx * yi]; k < f->fcp[x1 + f->pnx * yi + 1]; k++) {
						int q = f->ids[k];
						if (q == i) continue;
						float dx = pos[2 * q] - px, dy = pos[2 * q + 1] - py;
						float d2 = dx * dx + dy * dy;
						if (d2 > min_dist_sq || d2 == 0) continue;
						float d = sqrtf(d2);
						float sc = .5f * (min_dist - d) / d;
						ax -= dx * sc;
						ay -= dy * sc;
						if (with_colors) {
							c0 += .001f * ((c0 + f->pcol[3 * q]) / 2.f - c0);
							c1 += .001f * ((c1 + f->pcol[3 * q + 1]) / 2.f - c1);
							c2 += .001f * ((c2 + f->pcol[3 * q + 2]) / 2.f - c2);
						}
					}
			next[2 * i] = ax; next[2 * i + 1] = ay;
			ncol[3 * i] = c0; ncol[3 * i + 1] = c1; ncol[3 * i + 2] = c2;
		}
		memcpy(f->ppos, next, 4 * n2);
		if (with_colors) memcpy(f->pcol, ncol, 4 * n3);
	}
	free(next); free(ncol);
}
