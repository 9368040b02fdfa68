/* TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
 *
 * Plain-C CPU restatement ("port") of the reference Eulerian smoke solver,
 *   /root/reference/eulerian_fluid/src/fluid.h  (class Fluid).
 * Pinned bit-for-bit against the reference itself (oracle/_ref/libref_euler.so)
 * in tests/test_oracle_pin.py and against tests/golden/ fixtures.
 * Layout is the reference's: ix(i,j)=i*num_y+j, y fastest (fluid.h:94-96).
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared.
 */
#include <stdlib.h>
#include <string.h>
#include <time.h>
#if defined(__x86_64__) || defined(__i386__)
#include <xmmintrin.h>
#endif

enum { U_FIELD = 0, V_FIELD = 1, S_FIELD = 2 }; /* fluid.h:33-37 */

typedef struct {
	int nx, ny, ncells;
	float density, h, h1;
	float *u, *v, *new_u, *new_v, *pressure, *m, *new_m;
	unsigned char* solid; /* the reference's bool[] */
} Euler;

#define IX(i, j) ((i) * f->ny + (j))

/* ctor, fluid.h:41-69: two border cells added to each axis */
void* orceuler_create(int x, int y, float density, float h) {
	Euler* f = (Euler*)calloc(1, sizeof(Euler));
	f->nx = 2 + x; f->ny = 2 + y; f->ncells = f->nx * f->ny;
	f->density = density; f->h = h; f->h1 = 1 / h;
	size_t C = (size_t)f->ncells;
	f->u = calloc(C, 4); f->v = calloc(C, 4); f->new_u = calloc(C, 4); f->new_v = calloc(C, 4);
	f->pressure = calloc(C, 4); f->m = calloc(C, 4); f->new_m = calloc(C, 4);
	f->solid = calloc(C, 1);
	for (size_t c = 0; c < C; c++) f->m[c] = 1.f; /* :63-68 */
	return f;
}
void orceuler_destroy(void* h) {
	Euler* f = (Euler*)h;
	free(f->u); free(f->v); free(f->new_u); free(f->new_v); free(f->pressure); free(f->m); free(f->new_m); free(f->solid);
	free(f);
}
void orceuler_dims(void* h, int* ints, float* floats) {
	Euler* f = (Euler*)h;
	ints[0] = f->nx; ints[1] = f->ny; ints[2] = f->ncells;
	floats[0] = f->density; floats[1] = f->h; floats[2] = f->h1;
}
void* orceuler_ptr(void* h, const char* name) {
	Euler* f = (Euler*)h;
#define FIELD(n) if (!strcmp(name, #n)) return (void*)f->n;
	FIELD(u) FIELD(v) FIELD(new_u) FIELD(new_v) FIELD(pressure) FIELD(solid) FIELD(m) FIELD(new_m)
#undef FIELD
	return NULL;
}

/* solveIncompressibility, fluid.h:98-139 (omega literal 1.9, integer neighbour count) */
void orceuler_solve(void* h, int iters, float dt) {
	Euler* f = (Euler*)h;
	float cp = f->density * f->h / dt;
	for (int c = 0; c < f->ncells; c++) f->pressure[c] = 0.f;
	for (int it = 0; it < iters; it++)
		for (int i = 1; i < f->nx - 1; i++)
			for (int j = 1; j < f->ny - 1; j++) {
				if (f->solid[IX(i, j)]) continue;
				int sx0 = !f->solid[IX(i - 1, j)], sx1 = !f->solid[IX(i + 1, j)];
				int sy0 = !f->solid[IX(i, j - 1)], sy1 = !f->solid[IX(i, j + 1)];
				int s = sx0 + sx1 + sy0 + sy1;
				if (s == 0) continue;
				float div = f->u[IX(i + 1, j)] - f->u[IX(i, j)] + f->v[IX(i, j + 1)] - f->v[IX(i, j)];
				float p = -div / s;
				p *= 1.9f;
				f->pressure[IX(i, j)] += cp * p;
				f->u[IX(i, j)] -= sx0 * p;
				f->u[IX(i + 1, j)] += sx1 * p;
				f->v[IX(i, j)] -= sy0 * p;
				f->v[IX(i, j + 1)] += sy1 * p;
			}
}

/* extrapolate, fluid.h:142-151 */
void orceuler_extrapolate(void* h) {
	Euler* f = (Euler*)h;
	for (int i = 0; i < f->nx; i++) {
		f->u[IX(i, 0)] = f->u[IX(i, 1)];
		f->u[IX(i, f->ny - 1)] = f->u[IX(i, f->ny - 2)];
	}
	for (int j = 0; j < f->ny; j++) {
		f->v[IX(0, j)] = f->v[IX(1, j)];
		f->v[IX(f->nx - 1, j)] = f->v[IX(f->nx - 2, j)];
	}
}

static float minf(float a, float b) { return a < b ? a : b; }
static float maxf(float a, float b) { return a > b ? a : b; }
static int mini(int a, int b) { return a < b ? a : b; }

/* sampleField, fluid.h:153-188 */
float orceuler_sample(void* h, float x, float y, int field) {
	Euler* f = (Euler*)h;
	x = maxf(f->h, minf(x, f->nx * f->h));
	y = maxf(f->h, minf(y, f->ny * f->h));
	float dx = 0.f, dy = 0.f, h2 = f->h / 2;
	const float* a = NULL;
	switch (field) {
		case U_FIELD: a = f->u; dy = h2; break;
		case V_FIELD: a = f->v; dx = h2; break;
		case S_FIELD: a = f->m; dx = h2; dy = h2; break;
	}
	if (!a) return 0.f;
	int x0 = mini((int)(f->h1 * (x - dx)), f->nx - 1);
	int y0 = mini((int)(f->h1 * (y - dy)), f->ny - 1);
	int x1 = mini(x0 + 1, f->nx - 1), y1 = mini(y0 + 1, f->ny - 1);
	float tx = f->h1 * ((x - dx) - f->h * x0);
	float ty = f->h1 * ((y - dy) - f->h * y0);
	float sx = 1 - tx, sy = 1 - ty;
	return sx * sy * a[IX(x0, y0)] + sx * ty * a[IX(x0, y1)] + tx * sy * a[IX(x1, y0)] + tx * ty * a[IX(x1, y1)];
}

/* advectVel, fluid.h:210-239 with avgU :191-198 and avgV :201-208 */
void orceuler_advect_vel(void* h, float dt) {
	Euler* f = (Euler*)h;
	size_t bytes = 4 * (size_t)f->ncells;
	memcpy(f->new_u, f->u, bytes);
	memcpy(f->new_v, f->v, bytes);
	float h2 = f->h / 2;
	for (int i = 1; i < f->nx; i++)
		for (int j = 1; j < f->ny; j++) {
			if (f->solid[IX(i, j)]) continue;
			if (!f->solid[IX(i - 1, j)] && j < f->ny - 1) {
				float x = f->h * i, y = h2 + f->h * j;
				float u0 = f->u[IX(i, j)];
				float v0 = (f->v[IX(i, j)] + f->v[IX(i, j + 1)] + f->v[IX(i - 1, j)] + f->v[IX(i - 1, j + 1)]) / 4;
				f->new_u[IX(i, j)] = orceuler_sample(h, x - dt * u0, y - dt * v0, U_FIELD);
			}
			if (!f->solid[IX(i, j - 1)] && i < f->nx - 1) {
				float x = h2 + f->h * i, y = f->h * j;
				float u0 = (f->u[IX(i, j)] + f->u[IX(i, j - 1)] + f->u[IX(i + 1, j)] + f->u[IX(i + 1, j - 1)]) / 4;
				float v0 = f->v[IX(i, j)];
				f->new_v[IX(i, j)] = orceuler_sample(h, x - dt * u0, y - dt * v0, V_FIELD);
			}
		}
	memcpy(f->u, f->new_u, bytes);
	memcpy(f->v, f->new_v, bytes);
}

/* advectSmoke, fluid.h:241-259 */
void orceuler_advect_smoke(void* h, float dt) {
	Euler* f = (Euler*)h;
	size_t bytes = 4 * (size_t)f->ncells;
	memcpy(f->new_m, f->m, bytes);
	float h2 = f->h / 2;
	for (int i = 1; i < f->nx - 1; i++)
		for (int j = 1; j < f->ny - 1; j++) {
			if (f->solid[IX(i, j)]) continue;
			float u0 = (f->u[IX(i, j)] + f->u[IX(i + 1, j)]) / 2;
			float v0 = (f->v[IX(i, j)] + f->v[IX(i, j + 1)]) / 2;
			float x = h2 + f->h * i - dt * u0;
			float y = h2 + f->h * j - dt * v0;
			f->new_m[IX(i, j)] = orceuler_sample(h, x, y, S_FIELD);
		}
	memcpy(f->m, f->new_m, bytes);
}

/* FluidUI::setObstacle, eulerian_fluid/src/fluid_ui.h:24-62.  The UI shell owns this function (it is not a member of
 * class Fluid and its header needs the windowing engine, so it cannot be compiled into oracle/_ref): parity for it is
 * anchored on this restatement + the independent numpy one in tests/.  (vx, vy) = (x - previous x) / dt when the
 * caller passes reset = false (:39-43), else 0.  int arithmetic like the reference ("int is used for overflow reasons"). */
void orceuler_set_obstacle(void* h, int x, int y, int r, float vx, float vy) {
	Euler* f = (Euler*)h;
	memset(f->solid, 0, (size_t)f->ncells);                       /* :25 */
	for (int j = 0; j < f->ny; j++) f->solid[IX(0, j)] = 1;       /* left wall :28-30 */
	for (int i = 0; i < f->nx; i++) {                             /* top, bottom :33-36 */
		f->solid[IX(i, 0)] = 1;
		f->solid[IX(i, f->ny - 1)] = 1;
	}
	for (int i = 1; i < f->nx - 1; i++) {                         /* disc :49-60 */
		int dx = i - x;
		for (int j = 1; j < f->ny - 1; j++) {
			int dy = j - y;
			if (dx * dx + dy * dy < r * r) {
				f->solid[IX(i, j)] = 1;
				f->m[IX(i, j)] = 1.f;
				f->u[IX(i, j)] = vx;
				f->v[IX(i, j)] = vy;
			}
		}
	}
}

/* SURVEY Q17 timing aid (see oracle/ref_harness_euler.cpp): MXCSR FTZ|DAZ of the calling thread; returns the old setting. */
int orceuler_set_ftz_daz(int on) {
#if defined(__x86_64__) || defined(__i386__)
	unsigned csr = _mm_getcsr();
	int was = (csr & 0x8040u) == 0x8040u;
	_mm_setcsr(on ? (csr | 0x8040u) : (csr & ~0x8040u));
	return was;
#else
	(void)on;
	return -1;
#endif
}

/* caller sequence, eulerian_fluid/src/fluid_ui.h:107-111 */
void orceuler_step(void* h, int iters, float dt, int n_steps, double* phase_ns) {
	for (int k = 0; k < n_steps; k++) {
		struct timespec a, b, c;
		clock_gettime(CLOCK_MONOTONIC, &a);
		orceuler_solve(h, iters, dt);
		clock_gettime(CLOCK_MONOTONIC, &b);
		orceuler_extrapolate(h);
		orceuler_advect_vel(h, dt);
		orceuler_advect_smoke(h, dt);
		clock_gettime(CLOCK_MONOTONIC, &c);
		if (phase_ns) {
			phase_ns[0] += 1e9 * (double)(b.tv_sec - a.tv_sec) + (double)(b.tv_nsec - a.tv_nsec);
			phase_ns[1] += 1e9 * (double)(c.tv_sec - b.tv_sec) + (double)(c.tv_nsec - b.tv_nsec);
		}
	}
}
