/* TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
 *
 * Plain-C CPU restatement ("port") of the reference 3D Eulerian smoke solver,
 *   /root/reference/eulerian_fluid/src/fluid3d.h  (struct Fluid3D; no reference target compiles it).
 * Pinned bit-for-bit against the reference header itself (oracle/_ref/libref_fluid3d.so) in tests/test_oracle_pin.py and
 * against tests/golden/ fixtures.  Layout is the reference's: ix(i,j,k) = i + num_x*j + num_y*num_x*k (fluid3d.h:131-133).
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared.
 */
#include <stdlib.h>
#include <string.h>

enum { U_FIELD = 0, V_FIELD = 1, W_FIELD = 2, S_FIELD = 3 }; /* fluid3d.h:31-36 */

typedef struct {
	int nx, ny, nz, ncells;
	float density, h;
	float *u, *v, *w, *new_u, *new_v, *new_w, *pressure, *m, *new_m;
	unsigned char* solid; /* the reference's bool[] */
} F3;

#define IX(i, j, k) ((i) + f->nx * (j) + f->ny * f->nx * (k))
static float fminr(float a, float b) { return a < b ? a : b; } /* the header's min/max templates, :8-16 */
static float fmaxr(float a, float b) { return a > b ? a : b; }
static int iminr(int a, int b) { return a < b ? a : b; }

/* ctor, fluid3d.h:40-69: two border cells added to each axis, u=v=w=0, solid=false, m=1 */
void* orcfluid3d_create(int x, int y, int z, float density, float h) {
	F3* f = (F3*)calloc(1, sizeof(F3));
	f->nx = 2 + x; f->ny = 2 + y; f->nz = 2 + z; f->ncells = f->nx * f->ny * f->nz;
	f->density = density; f->h = h;
	size_t C = (size_t)f->ncells;
	f->u = calloc(C, 4); f->v = calloc(C, 4); f->w = calloc(C, 4); f->new_u = calloc(C, 4); f->new_v = calloc(C, 4); f->new_w = calloc(C, 4);
	f->pressure = calloc(C, 4); f->m = calloc(C, 4); f->new_m = calloc(C, 4);
	f->solid = calloc(C, 1);
	for (size_t c = 0; c < C; c++) f->m[c] = 1.f;
	return f;
}
void orcfluid3d_destroy(void* h) {
	F3* f = (F3*)h;
	free(f->u); free(f->v); free(f->w); free(f->new_u); free(f->new_v); free(f->new_w); free(f->pressure); free(f->m); free(f->new_m); free(f->solid);
	free(f);
}
void orcfluid3d_dims(void* h, int* ints, float* floats) {
	F3* f = (F3*)h;
	ints[0] = f->nx; ints[1] = f->ny; ints[2] = f->nz; ints[3] = f->ncells;
	floats[0] = f->density; floats[1] = f->h;
}
void* orcfluid3d_ptr(void* h, const char* name) {
	F3* f = (F3*)h;
#define FIELD(n) if (!strcmp(name, #n)) return (void*)f->n;
	FIELD(u) FIELD(v) FIELD(w) FIELD(new_u) FIELD(new_v) FIELD(new_w) FIELD(pressure) FIELD(solid) FIELD(m) FIELD(new_m)
#undef FIELD
	return NULL;
}

/* solveIncompressibility, fluid3d.h:135-182: i outer, j, k inner; integer neighbour count; omega literal 1.9 */
void orcfluid3d_solve(void* h, int iters, float dt) {
	F3* f = (F3*)h;
	float cp = f->density * f->h / dt;
	for (int c = 0; c < f->ncells; c++) f->pressure[c] = 0.f;
	for (int it = 0; it < iters; it++)
		for (int i = 1; i < f->nx - 1; i++)
			for (int j = 1; j < f->ny - 1; j++)
				for (int k = 1; k < f->nz - 1; k++) {
					if (f->solid[IX(i, j, k)]) continue;
					int sx0 = !f->solid[IX(i - 1, j, k)], sx1 = !f->solid[IX(i + 1, j, k)], sy0 = !f->solid[IX(i, j - 1, k)];
					int sy1 = !f->solid[IX(i, j + 1, k)], sz0 = !f->solid[IX(i, j, k - 1)], sz1 = !f->solid[IX(i, j, k + 1)];
					int s = sx0 + sx1 + sy0 + sy1 + sz0 + sz1;
					if (s == 0) continue;
					float div = f->u[IX(i + 1, j, k)] - f->u[IX(i, j, k)] + f->v[IX(i, j + 1, k)] - f->v[IX(i, j, k)] + f->w[IX(i, j, k + 1)] - f->w[IX(i, j, k)];
					float p = -div / s;
					p *= 1.9f;
					f->pressure[IX(i, j, k)] += cp * p;
					f->u[IX(i, j, k)] -= sx0 * p; f->u[IX(i + 1, j, k)] += sx1 * p;
					f->v[IX(i, j, k)] -= sy0 * p; f->v[IX(i, j + 1, k)] += sy1 * p;
					f->w[IX(i, j, k)] -= sz0 * p; f->w[IX(i, j, k + 1)] += sz1 * p;
				}
}

/* extrapolate, fluid3d.h:185-211: three border loops one after the other */
void orcfluid3d_extrapolate(void* h) {
	F3* f = (F3*)h;
	for (int i = 0; i < f->nx; i++)
		for (int j = 0; j < f->ny; j++) {
			f->u[IX(i, j, 0)] = f->u[IX(i, j, 1)]; f->u[IX(i, j, f->nz - 1)] = f->u[IX(i, j, f->nz - 2)];
			f->v[IX(i, j, 0)] = f->v[IX(i, j, 1)]; f->v[IX(i, j, f->nz - 1)] = f->v[IX(i, j, f->nz - 2)];
		}
	for (int j = 0; j < f->ny; j++)
		for (int k = 0; k < f->nz; k++) {
			f->v[IX(0, j, k)] = f->v[IX(1, j, k)]; f->v[IX(f->nx - 1, j, k)] = f->v[IX(f->nx - 2, j, k)];
			f->w[IX(0, j, k)] = f->w[IX(1, j, k)]; f->w[IX(f->nx - 1, j, k)] = f->w[IX(f->nx - 2, j, k)];
		}
	for (int k = 0; k < f->nz; k++)
		for (int i = 0; i < f->nx; i++) {
			f->w[IX(i, 0, k)] = f->w[IX(i, 1, k)]; f->w[IX(i, f->ny - 1, k)] = f->w[IX(i, f->ny - 2, k)];
			f->u[IX(i, 0, k)] = f->u[IX(i, 1, k)]; f->u[IX(i, f->ny - 1, k)] = f->u[IX(i, f->ny - 2, k)];
		}
}

/* sampleField, fluid3d.h:213-264 (S_FIELD is offset in x and y only, :247) */
static float sample(const F3* f, float x, float y, float z, int field) {
	float h = f->h, h1 = 1 / h, h2 = h / 2;
	x = fmaxr(h, fminr(x, f->nx * h));
	y = fmaxr(h, fminr(y, f->ny * h));
	z = fmaxr(h, fminr(z, f->nz * h));
	float dx = 0.f, dy = 0.f, dz = 0.f;
	const float* a = NULL;
	switch (field) {
		case U_FIELD: a = f->u; dy = h2; break;
		case V_FIELD: a = f->v; dx = h2; break;
		case W_FIELD: a = f->w; dz = h2; break;
		case S_FIELD: a = f->m; dx = h2; dy = h2; break;
	}
	if (!a) return 0.f;
	int x0 = iminr((int)(h1 * (x - dx)), f->nx - 1), y0 = iminr((int)(h1 * (y - dy)), f->ny - 1), z0 = iminr((int)(h1 * (z - dz)), f->nz - 1);
	int x1 = iminr(x0 + 1, f->nx - 1), y1 = iminr(y0 + 1, f->ny - 1), z1 = iminr(z0 + 1, f->nz - 1);
	float tx = h1 * ((x - dx) - h * x0), ty = h1 * ((y - dy) - h * y0), tz = h1 * ((z - dz) - h * z0);
	float sx = 1 - tx, sy = 1 - ty, sz = 1 - tz;
	return sx * sy * sz * a[IX(x0, y0, z0)] + sx * sy * tz * a[IX(x0, y0, z1)] + sx * ty * sz * a[IX(x0, y1, z0)] + sx * ty * tz * a[IX(x0, y1, z1)] +
	       tx * sy * sz * a[IX(x1, y0, z0)] + tx * sy * tz * a[IX(x1, y0, z1)] + tx * ty * sz * a[IX(x1, y1, z0)] + tx * ty * tz * a[IX(x1, y1, z1)];
}
float orcfluid3d_sample(void* h, float x, float y, float z, int field) { return sample((const F3*)h, x, y, z, field); }

static float avgU(const F3* f, int i, int j, int k) { /* :266-277 */
	return (f->u[IX(i, j, k)] + f->u[IX(i, j, k - 1)] + f->u[IX(i, j - 1, k)] + f->u[IX(i, j - 1, k - 1)] + f->u[IX(i + 1, j, k)] + f->u[IX(i + 1, j, k - 1)] +
	        f->u[IX(i + 1, j - 1, k)] + f->u[IX(i + 1, j - 1, k - 1)]) / 8;
}
static float avgV(const F3* f, int i, int j, int k) { /* :279-290 */
	return (f->v[IX(i, j, k)] + f->v[IX(i, j, k - 1)] + f->v[IX(i - 1, j, k)] + f->v[IX(i - 1, j, k - 1)] + f->v[IX(i, j + 1, k)] + f->v[IX(i, j + 1, k - 1)] +
	        f->v[IX(i - 1, j + 1, k)] + f->v[IX(i - 1, j + 1, k - 1)]) / 8;
}
static float avgW(const F3* f, int i, int j, int k) { /* :292-303 */
	return (f->w[IX(i, j, k)] + f->w[IX(i, j - 1, k)] + f->w[IX(i - 1, j, k)] + f->w[IX(i - 1, j - 1, k)] + f->w[IX(i, j, k + 1)] + f->w[IX(i, j - 1, k + 1)] +
	        f->w[IX(i - 1, j, k + 1)] + f->w[IX(i - 1, j - 1, k + 1)]) / 8;
}

/* advectVel, fluid3d.h:305-346.  The k loop is bounded by num_y (:311); for num_z < num_y the reference runs out of bounds,
 * this port stops at min(num_y, num_z) */
void orcfluid3d_advect_vel(void* hh, float dt) {
	F3* f = (F3*)hh;
	size_t bytes = 4 * (size_t)f->ncells;
	memcpy(f->new_u, f->u, bytes); memcpy(f->new_v, f->v, bytes); memcpy(f->new_w, f->w, bytes);
	float h = f->h, h2 = h / 2;
	int kmax = iminr(f->ny, f->nz);
	for (int i = 1; i < f->nx; i++)
		for (int j = 1; j < f->ny; j++)
			for (int k = 1; k < kmax; k++) {
				if (f->solid[IX(i, j, k)]) continue;
				if (!f->solid[IX(i - 1, j, k)] && j < f->ny - 1 && k < f->nz - 1) {
					float x = h * i, y = h2 + h * j, z = h2 + h * k;
					float u0 = f->u[IX(i, j, k)], v0 = avgV(f, i, j, k), w0 = avgW(f, i, j, k);
					f->new_u[IX(i, j, k)] = sample(f, x - dt * u0, y - dt * v0, z - dt * w0, U_FIELD);
				}
				if (!f->solid[IX(i, j - 1, k)] && k < f->nz - 1 && i < f->nx - 1) {
					float x = h2 + h * i, y = h * j, z = h2 + h * k;
					float u0 = avgU(f, i, j, k), v0 = f->v[IX(i, j, k)], w0 = avgW(f, i, j, k);
					f->new_v[IX(i, j, k)] = sample(f, x - dt * u0, y - dt * v0, z - dt * w0, V_FIELD);
				}
				if (!f->solid[IX(i, j, k - 1)] && i < f->nx - 1 && j < f->ny - 1) {
					float x = h2 + h * i, y = h2 + h * j, z = h * k;
					float u0 = avgU(f, i, j, k), v0 = avgV(f, i, j, k), w0 = f->w[IX(i, j, k)];
					f->new_w[IX(i, j, k)] = sample(f, x - dt * u0, y - dt * v0, z - dt * w0, W_FIELD);
				}
			}
	memcpy(f->u, f->new_u, bytes); memcpy(f->v, f->new_v, bytes); memcpy(f->w, f->new_w, bytes);
}

/* advectSmoke, fluid3d.h:348-367 */
void orcfluid3d_advect_smoke(void* hh, float dt) {
	F3* f = (F3*)hh;
	size_t bytes = 4 * (size_t)f->ncells;
	memcpy(f->new_m, f->m, bytes);
	float h = f->h, h2 = h / 2;
	for (int i = 1; i < f->nx - 1; i++)
		for (int j = 1; j < f->ny - 1; j++)
			for (int k = 1; k < f->nz - 1; k++) {
				if (f->solid[IX(i, j, k)]) continue;
				float u0 = (f->u[IX(i, j, k)] + f->u[IX(i + 1, j, k)]) / 2;
				float v0 = (f->v[IX(i, j, k)] + f->v[IX(i, j + 1, k)]) / 2;
				float w0 = (f->w[IX(i, j, k)] + f->w[IX(i, j, k + 1)]) / 2;
				float x = h2 + h * i - dt * u0, y = h2 + h * j - dt * v0, z = h2 + h * k - dt * w0;
				f->new_m[IX(i, j, k)] = sample(f, x, y, z, S_FIELD);
			}
	memcpy(f->m, f->new_m, bytes);
}

void orcfluid3d_step(void* h, int iters, float dt, int n_steps) {
	for (int k = 0; k < n_steps; k++) {
		orcfluid3d_solve(h, iters, dt);
		orcfluid3d_extrapolate(h);
		orcfluid3d_advect_vel(h, dt);
		orcfluid3d_advect_smoke(h, dt);
	}
}
