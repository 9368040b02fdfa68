// Headless driver: the reference demos with olc::PixelGameEngine stripped out, written once as
// `template <class Sim>` so the SAME scene-building and stepping code runs either
//   * flipgpu::FlipFluidGPU / flipgpu::FluidGPU  (this repo; default build -> headless_gpu), or
//   * the reference's FlipFluid / Fluid           (-DWITH_REFERENCE, built by oracle/Makefile as test infrastructure).
// Scene code follows flip_fluid/src/main.cpp:33-126,162-165 and eulerian_fluid/src/fluid_ui.h:25-111.
//
//   headless flip  default      <steps>
//   headless flip  synthetic N  <steps> [obstacle]
//   headless euler X Y ITERS    <steps>
// Optional trailing  --ppm FILE  writes the frame the UI would draw after the last step (main.cpp:171-196 /
// fluid_ui.h:126-151) as a binary PPM: the reference's only "test" is looking at that frame.
// Prints one JSON line: sizes, wall time, and checksums of the state the UI would draw.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#ifdef WITH_REFERENCE
#include "flip_fluid.h"  // -I/root/reference/flip_fluid/src (unmodified)
using FlipSimT = FlipFluid;
#else
#include "flip_fluid_gpu.hpp"
using FlipSimT = flipgpu::FlipFluidGPU;
#endif

template <class T, class = void> struct has_mark_dirty : std::false_type {};
template <class T> struct has_mark_dirty<T, std::void_t<decltype(std::declval<T&>().mark_dirty())>> : std::true_type {};
template <class Sim> void host_wrote(Sim& s) { if constexpr (has_mark_dirty<Sim>::value) s.mark_dirty(); }
template <class Sim> void host_reads(Sim& s) { if constexpr (has_mark_dirty<Sim>::value) s.pull_render_state(); }


static const char* g_ppm_path = nullptr;

static void write_ppm(const char* path, int w, int h, const std::vector<unsigned char>& rgb) {
	FILE* fp = std::fopen(path, "wb");
	if (!fp) { std::perror(path); return; }
	std::fprintf(fp, "P6\n%d %d\n255\n", w, h);
	std::fwrite(rgb.data(), 1, rgb.size(), fp);
	std::fclose(fp);
}
static unsigned char to_byte(float c) { c = c < 0 ? 0 : (c > 1 ? 1 : c); return (unsigned char)(c * 255.f + .5f); }

// the frame of flip_fluid/src/main.cpp:171-196 at one pixel per 1/px_per_cell of a cell: density cells (cell_color) as the
// background, then every particle as a dot in its particle_color, then the obstacle outline in red
template <class Sim> void dump_flip_frame(Sim& f, const char* path, float ox, float oy, float orad) {
	const int px = 4, w = f.f_num_x * px, h = f.f_num_y * px;
	std::vector<unsigned char> img((size_t)w * h * 3);
	for (int y = 0; y < h; y++)
		for (int x = 0; x < w; x++) {
			int c = f.fIX(x / px, y / px);
			for (int k = 0; k < 3; k++) img[3 * ((size_t)y * w + x) + k] = to_byte(f.cell_color[3 * c + k] * .35f);
		}
	const float scale = px / f.h;
	for (int i = 0; i < f.num_particles; i++) {
		int x = (int)(f.particle_pos[2 * i] * scale), y = (int)(f.particle_pos[2 * i + 1] * scale);
		if (x < 0 || x >= w || y < 0 || y >= h) continue;
		for (int k = 0; k < 3; k++) img[3 * ((size_t)y * w + x) + k] = to_byte(f.particle_color[3 * i + k]);
	}
	if (orad > 0)
		for (int a = 0; a < 720; a++) {
			int x = (int)((ox + orad * std::cos(a * 3.14159265f / 360)) * scale), y = (int)((oy + orad * std::sin(a * 3.14159265f / 360)) * scale);
			if (x >= 0 && x < w && y >= 0 && y < h) { img[3 * ((size_t)y * w + x)] = 255; img[3 * ((size_t)y * w + x) + 1] = 0; img[3 * ((size_t)y * w + x) + 2] = 0; }
		}
	write_ppm(path, w, h, img);
}

struct FlipScene {
	float width, height, spacing, radius, dt;
	float obstacle_x = 0, obstacle_y = 0, obstacle_radius = 0;
	int iters = 50;
	bool obstacle = false;
};

// FluidUI::setObstacle with reset=true, main.cpp:95-126 (Q10: loops stop at n-2)
template <class Sim> void set_obstacle(Sim& f, float x, float y, float rad) {
	for (int i = 1; i < f.f_num_x - 2; i++)
		for (int j = 1; j < f.f_num_y - 2; j++) {
			int c = f.fIX(i, j);
			f.s[c] = 1;
			float dx = f.h * (.5f + i) - x, dy = f.h * (.5f + j) - y;
			if (dx * dx + dy * dy < rad * rad) {
				f.s[c] = 0;
				f.u[c] = 0; f.u[c + f.fIX(1, 0)] = 0; f.v[c] = 0; f.v[c + f.fIX(0, 1)] = 0;
			}
		}
}

template <class Sim> int run_flip(const FlipScene& sc, int steps) {
	// OnUserCreate, main.cpp:46-79
	float h = sc.spacing, r = sc.radius;
	float dx = 2 * r;
	float dy = std::sqrt(3) / 2 * dx;
	float rel_water_height = .8f, rel_water_width = .6f;
	int num_x = (rel_water_width * sc.width - 2 * h - 2 * r) / dx;
	int num_y = (rel_water_height * sc.height - 2 * h - 2 * r) / dy;
	int max_particles = num_x * num_y;
	Sim* fluid = new Sim(1000.f, sc.width, sc.height, h, r, max_particles);
#ifdef WITH_REFERENCE
	{  // Q4: the reference leaves its state uninitialised; the contract is all-zero
		size_t C = fluid->f_num_cells;
		for (float* a : {fluid->u, fluid->v, fluid->du, fluid->dv, fluid->prev_u, fluid->prev_v, fluid->p, fluid->s, fluid->particle_density})
			std::memset(a, 0, 4 * C);
		std::memset(fluid->cell_color, 0, 12 * C);
		std::memset(fluid->cell_type, 0, 4 * C);
		std::memset(fluid->particle_vel, 0, 8 * (size_t)max_particles);
		for (int i = 0; i < max_particles; i++) fluid->particle_color[3 * i] = fluid->particle_color[3 * i + 1] = 0;
	}
#endif
#ifndef WITH_REFERENCE
	if (std::getenv("FLIPGPU_EXACT")) flipgpu::check(flip_set_exact_order(fluid->handle(), 1), "flip_set_exact_order");  // parity runs
#endif
	fluid->num_particles = num_x * num_y;
	int p = 0;
	for (int i = 0; i < num_x; i++)
		for (int j = 0; j < num_y; j++) {
			fluid->particle_pos[p++] = h + r + dx * i + (j % 2 == 0 ? 0 : r);
			fluid->particle_pos[p++] = h + r + dy * j;
		}
	for (int i = 0; i < fluid->f_num_x; i++)
		for (int j = 0; j < fluid->f_num_y; j++) {
			float s = 1;
			if (i == 0 || i == fluid->f_num_x - 1 || j == 0 || j == fluid->f_num_y - 1) s = 0;
			fluid->s[fluid->fIX(i, j)] = s;
		}
	if (sc.obstacle) set_obstacle(*fluid, sc.obstacle_x, sc.obstacle_y, sc.obstacle_radius);
	host_wrote(*fluid);
	// the fixed-timestep loop of OnUserUpdate, main.cpp:160-168 (Q18: headless just loops)
	auto t0 = std::chrono::steady_clock::now();
	for (int k = 0; k < steps; k++)
		fluid->simulate(sc.dt, 9.81f, .9f, sc.iters, 2, 1.9f, true, true, sc.obstacle_x, sc.obstacle_y, 0.f, 0.f, sc.obstacle_radius);
	host_reads(*fluid);  // what the renderer draws: particle_pos, particle_color, cell_color (main.cpp:173-194)
	double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	if (g_ppm_path) dump_flip_frame(*fluid, g_ppm_path, sc.obstacle_x, sc.obstacle_y, sc.obstacle ? sc.obstacle_radius : 0.f);
	double sx = 0, sy = 0, sc_b = 0, cell = 0;
	for (int i = 0; i < fluid->num_particles; i++) {
		sx += fluid->particle_pos[2 * i]; sy += fluid->particle_pos[2 * i + 1]; sc_b += fluid->particle_color[3 * i + 2];
	}
	for (int i = 0; i < 3 * fluid->f_num_cells; i++) cell += fluid->cell_color[i];
	int P = fluid->num_particles;
	std::printf("{\"sim\": \"flip\", \"impl\": \"%s\", \"grid\": [%d, %d], \"particles\": %d, \"steps\": %d, \"seconds\": %.6f, "
	            "\"particle_steps_per_s\": %.6g, \"mean_x\": %.9g, \"mean_y\": %.9g, \"mean_blue\": %.9g, \"cell_color_sum\": %.9g}\n",
#ifdef WITH_REFERENCE
	            "reference",
#else
	            "flipgpu",
#endif
	            fluid->f_num_x, fluid->f_num_y, P, steps, secs, steps > 0 ? (double)P * steps / secs : 0.0, sx / P, sy / P, sc_b / P, cell);
	delete fluid;
	return 0;
}

#ifndef WITH_REFERENCE
// eulerian_fluid: Fluid's public arrays are raw pointers in the reference and vectors here; same indexing syntax
template <class Sim> int run_euler(int x, int y, int iters, int steps) {
	Sim* fluid = new Sim(x, y, 1000, 1 / 100.f);  // fluid_ui.h:69
	float in_vel = 2;
	for (int j = 0; j < fluid->getNumY(); j++) fluid->u[fluid->ix(1, j)] = in_vel;  // :72-75
	{  // setObstacle(numX/3, numY/2, numX/8), fluid_ui.h:25-62
		int ox = fluid->getNumX() / 3, oy = fluid->getNumY() / 2, r = fluid->getNumX() / 8;
		for (int c = 0; c < fluid->getNumCells(); c++) fluid->solid[c] = false;
		for (int j = 0; j < fluid->getNumY(); j++) fluid->solid[fluid->ix(0, j)] = true;
		for (int i = 0; i < fluid->getNumX(); i++) { fluid->solid[fluid->ix(i, 0)] = true; fluid->solid[fluid->ix(i, fluid->getNumY() - 1)] = true; }
		for (int i = 1; i < fluid->getNumX() - 1; i++) {
			int ddx = i - ox;
			for (int j = 1; j < fluid->getNumY() - 1; j++) {
				int ddy = j - oy;
				if (ddx * ddx + ddy * ddy < r * r) {
					fluid->solid[fluid->ix(i, j)] = true; fluid->m[fluid->ix(i, j)] = 1.f;
					fluid->u[fluid->ix(i, j)] = 0; fluid->v[fluid->ix(i, j)] = 0;
				}
			}
		}
	}
	int pipe_h = fluid->getNumY() / 6;  // :81-86
	int min_j = fluid->getNumY() / 2 - pipe_h / 2, max_j = fluid->getNumY() / 2 + pipe_h / 2;
	for (int j = min_j; j <= max_j; j++) fluid->m[fluid->ix(0, j)] = 0;
	host_wrote(*fluid);
	float dt = 1 / 60.f;
	auto t0 = std::chrono::steady_clock::now();
	for (int k = 0; k < steps; k++) {  // fluid_ui.h:107-111
		fluid->solveIncompressibility(iters, dt);
		fluid->extrapolate();
		fluid->advectVel(dt);
		fluid->advectSmoke(dt);
	}
	host_reads(*fluid);
	double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	if (g_ppm_path) {  // grayscale smoke, solid cells left black (fluid_ui.h:129-150), image x = i, y = j
		int w = fluid->getNumX(), h = fluid->getNumY();
		std::vector<unsigned char> img((size_t)w * h * 3, 0);
		for (int i = 0; i < w; i++)
			for (int j = 0; j < h; j++) {
				if (fluid->solid[fluid->ix(i, j)]) continue;
				unsigned char v = to_byte(fluid->m[fluid->ix(i, j)]);
				for (int k = 0; k < 3; k++) img[3 * ((size_t)j * w + i) + k] = v;
			}
		write_ppm(g_ppm_path, w, h, img);
	}
	double sm = 0, sp = 0;
	for (int c = 0; c < fluid->getNumCells(); c++) { sm += fluid->m[c]; sp += std::fabs(fluid->pressure[c]); }
	double updates = (double)(fluid->getNumX() - 2) * (fluid->getNumY() - 2) * iters * steps;
	std::printf("{\"sim\": \"euler\", \"impl\": \"flipgpu\", \"grid\": [%d, %d], \"iters\": %d, \"steps\": %d, \"seconds\": %.6f, "
	            "\"cell_updates_per_s\": %.6g, \"smoke_mean\": %.9g, \"abs_pressure_mean\": %.9g}\n",
	            fluid->getNumX(), fluid->getNumY(), iters, steps, secs, steps > 0 ? updates / secs : 0.0, sm / fluid->getNumCells(), sp / fluid->getNumCells());
	delete fluid;
	return 0;
}
#endif

int main(int argc, char** argv) {
	for (int i = 1; i + 1 < argc; i++)
		if (std::string(argv[i]) == "--ppm") { g_ppm_path = argv[i + 1]; argc = i; break; }
	if (argc < 4) {
		std::fprintf(stderr, "usage: %s flip default <steps> | flip synthetic N <steps> [obstacle] | euler X Y ITERS <steps>\n", argv[0]);
		return 2;
	}
	try {
		std::string what = argv[1];
		if (what == "flip") {
			FlipScene sc;
			int steps;
			if (std::string(argv[2]) == "default") {  // main.cpp:33-47 with the 640x480 window
				float sim_height = 3, c_scale = 480 / sim_height, sim_width = 640 / c_scale;
				sc.width = sim_width; sc.height = sim_height; sc.spacing = sim_height / 100; sc.radius = .3f * sc.spacing;
				sc.dt = 1 / 60.f; sc.obstacle = true; sc.obstacle_x = 3; sc.obstacle_y = 2; sc.obstacle_radius = .15f;
				steps = std::atoi(argv[3]);
			} else {  // SURVEY 8d rule for the synthetic configs
				if (argc < 5) return 2;
				int N = std::atoi(argv[3]);
				steps = std::atoi(argv[4]);
				sc.spacing = 3.f / N; sc.width = sc.height = (N - .5f) * sc.spacing; sc.radius = .19f * sc.spacing;
				sc.dt = (1 / 60.f) * (100.f / N);
				if (argc > 5) { sc.obstacle = true; sc.obstacle_x = .75f * sc.width; sc.obstacle_y = .67f * sc.height; sc.obstacle_radius = .05f * sc.width; sc.iters = 200; }
			}
			return run_flip<FlipSimT>(sc, steps);
		}
#ifndef WITH_REFERENCE
		if (what == "euler" && argc >= 6) return run_euler<flipgpu::FluidGPU>(std::atoi(argv[2]), std::atoi(argv[3]), std::atoi(argv[4]), std::atoi(argv[5]));
#endif
	} catch (const std::exception& e) {
		std::fprintf(stderr, "headless: %s\n", e.what());
		return 1;
	}
	return 2;
}
