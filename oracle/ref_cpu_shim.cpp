/*
 * ref_cpu_shim.cpp — TEST INFRASTRUCTURE ONLY (oracle/_ref/libref_cpu.so).
 *
 * Compiles the reference's OWN CPU implementation (cpu_launcher.cpp, taken from where it lies under
 * /root/reference via -DREF_CPU_SOURCE, never copied) and exposes its classes through a C entry point, so
 * that (a) the oracle restatement can be validated against the real reference at any resolution and for both
 * scene layouts, and (b) bench.py --impl reference / cpu_baseline.kind == "reference" can time the
 * reference's render loop on the GPU box's host cores. The only code written here is the launcher loop of
 * cpu_launcher.cpp:654-718 with W, H, the object order and the mesh path turned into arguments; ray
 * generation, intersection, traversal and shading are the reference's compiled code.
 */
#define main ref_cpu_launcher_main
#include REF_CPU_SOURCE
#undef main

#include <chrono>

namespace {
struct Built {
	Scene scene;
	TriangleMesh* mesh = nullptr;
};

float g_light[3] = {-10.f, 20.f, 40.f}; /* ref_cpu_set_light: Scene::L of the next scenes (cpu_launcher.cpp:650) */

/* scene_kind 2: BASELINE.json configs[0] / configs[3] — the six walls + the demo spheres of the commented lines
 *               cpu_launcher.cpp:668-672 (white diffuse, mirror, refractive shell = inner R 9 inside outer R 10), no mesh;
 *               ids as oracle/scenes.py:spheres_scene gives them (walls 0-5, then the four spheres)
 * scene_kind 0: cpu_launcher.cpp:673-685 (walls 0-5, cat 6, no rescale)
 * scene_kind 1: the object order and mesh transform of optimized.cu:684-726,804 (wall 0, cat 1, walls 2-6,
 *               rescale 0.6 / (0,-4,0)) built from cpu_launcher.cpp's classes. */
void build_scene(Built& b, const char* obj_path, int scene_kind) {
	Sphere* walls[6] = {
		new Sphere(Vector(0, 0, -1000), 940, Vector(0., 1., 0.)),
		new Sphere(Vector(0, -1000, 0), 990, Vector(0., 0., 1.)),
		new Sphere(Vector(0, 1000, 0), 940, Vector(1., 0., 0.)),
		new Sphere(Vector(-1000, 0, 0), 940, Vector(0., 1., 1.)),
		new Sphere(Vector(1000, 0, 0), 940, Vector(1., 1., 0.)),
		new Sphere(Vector(0, 0, 1000), 940, Vector(1., 0., 1.)),
	};
	TriangleMesh* mesh = nullptr;
	if (obj_path && obj_path[0]) {
		mesh = new TriangleMesh();
		mesh->readOBJ(obj_path);
		mesh->albedo = Vector(0.25, 0.25, 0.25);
		if (scene_kind == 1) {
			Vector offset(0.f, -4.f, 0.f);
			for (size_t i = 0; i < mesh->vertices.size(); i++) mesh->vertices[i] = mesh->vertices[i] * 0.6f + offset;
		}
		mesh->buildBVH(&(mesh->bvh), 0, mesh->indices.size());
	}
	b.mesh = mesh;
	b.scene.L = Vector(g_light[0], g_light[1], g_light[2]);
	if (scene_kind == 2) {
		for (int k = 0; k < 6; k++) b.scene.addObject(walls[k]);
		b.scene.addObject(new Sphere(Vector(0, 0, 0), 10, Vector(1., 1., 1.)));
		b.scene.addObject(new Sphere(Vector(-20, 0, 0), 10, Vector(0., 0., 0.), 1));
		b.scene.addObject(new Sphere(Vector(20, 0, 0), 9, Vector(0., 0., 0.), 0, 1, 1.5));
		b.scene.addObject(new Sphere(Vector(20, 0, 0), 10, Vector(0., 0., 0.), 0, 1.5, 1));
	} else if (scene_kind == 1) {
		b.scene.addObject(walls[0]);
		if (mesh) b.scene.addObject(mesh);
		for (int k = 1; k < 6; k++) b.scene.addObject(walls[k]);
	} else {
		for (int k = 0; k < 6; k++) b.scene.addObject(walls[k]);
		if (mesh) b.scene.addObject(mesh);
	}
}

void count_nodes(const BVH* n, int depth, int& nodes, int& leaves, int& max_depth, int& max_leaf) {
	nodes++;
	if (depth > max_depth) max_depth = depth;
	if (!n->left) {
		leaves++;
		if (n->triangle_end - n->triangle_start > max_leaf) max_leaf = n->triangle_end - n->triangle_start;
		return;
	}
	count_nodes(n->left, depth + 1, nodes, leaves, max_depth, max_leaf);
	count_nodes(n->right, depth + 1, nodes, leaves, max_depth, max_leaf);
}

/* pre-order dump in the 10-float layout of optimized.cu:512-534, from the pointer tree the reference built */
void dump_nodes(const BVH* n, float* arr, int& next, int idx) {
	float* a = arr + idx * 10;
	a[2] = n->bb.mn[0]; a[3] = n->bb.mn[1]; a[4] = n->bb.mn[2];
	a[5] = n->bb.mx[0]; a[6] = n->bb.mx[1]; a[7] = n->bb.mx[2];
	a[8] = n->triangle_start; a[9] = n->triangle_end;
	if (n->left) {
		int l = next++;
		a[0] = l;
		dump_nodes(n->left, arr, next, l);
		int r = next++;
		a[1] = r;
		dump_nodes(n->right, arr, next, r);
	} else {
		a[0] = -1; a[1] = -1;
	}
}
} // namespace

extern "C" {

/* Scene::L of the scenes built from now on (the light orbit of BASELINE.json configs[3]). */
void ref_cpu_set_light(float x, float y, float z) {
	g_light[0] = x;
	g_light[1] = y;
	g_light[2] = z;
}

/* Render with the reference's classes. rgb: H*W*3; obj_id: H*W primary-ray object id; P_out/N_out: H*W*3
 * primary hit point / normal as intersect_all returns them. Any output may be NULL. seconds = render loop only. */
int ref_cpu_render(const char* obj_path, int scene_kind, int W, int H, int num_rays, int num_bounce, int threads,
                   unsigned char* rgb, int* obj_id, float* P_out, float* N_out, double* seconds) {
	Built b;
	build_scene(b, obj_path, scene_kind);
	Scene& s = b.scene;
	if (threads > 0) omp_set_num_threads(threads);
	float alpha = PI / 3;
	Vector C(0, 0, 55);
	float z = -W / (2 * tan(alpha / 2));
	auto t0 = std::chrono::steady_clock::now();
	#pragma omp parallel for schedule(dynamic, 1)
	for (int i = 0; i < H; i++) {
		for (int j = 0; j < W; j++) {
			unsigned int seed = omp_get_thread_num();
			Vector u_center((float)j - (float)W / 2 + 0.5, (float)H / 2 - i - 0.5, z);
			Vector color_total(0, 0, 0);
			for (int t = 0; t < num_rays; t++) {
				float sigma = 0;
				float r1 = uniform(seed);
				float r2 = uniform(seed);
				Vector u = u_center + Vector(sigma * sqrt(-2 * log(r1)) * cos(2 * PI * r2), sigma * sqrt(-2 * log(r1)) * sin(2 * PI * r2), 0);
				u.normalize();
				Ray r(C, u);
				color_total = color_total + s.getColor(r, num_bounce);
			}
			Vector color_avg = color_total / num_rays;
			if (rgb) {
				rgb[(i * W + j) * 3 + 0] = std::min(std::pow(color_avg[0], 1. / 2.2), 255.);
				rgb[(i * W + j) * 3 + 1] = std::min(std::pow(color_avg[1], 1. / 2.2), 255.);
				rgb[(i * W + j) * 3 + 2] = std::min(std::pow(color_avg[2], 1. / 2.2), 255.);
			}
		}
	}
	auto t1 = std::chrono::steady_clock::now();
	if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
	if (obj_id || P_out || N_out) {
		#pragma omp parallel for schedule(dynamic, 1)
		for (int i = 0; i < H; i++) {
			for (int j = 0; j < W; j++) {
				Vector u((float)j - (float)W / 2 + 0.5, (float)H / 2 - i - 0.5, z);
				u.normalize();
				Ray r(C, u);
				Vector P, N;
				int id = -1;
				s.intersect_all(r, P, N, id);
				size_t px = (size_t)i * W + j;
				if (obj_id) obj_id[px] = id;
				if (P_out) { P_out[px * 3] = P[0]; P_out[px * 3 + 1] = P[1]; P_out[px * 3 + 2] = P[2]; }
				if (N_out) { N_out[px * 3] = N[0]; N_out[px * 3 + 1] = N[1]; N_out[px * 3 + 2] = N[2]; }
			}
		}
	}
	return 0;
}

/* Mesh + BVH as the reference's loader and builder produce them (for the loader/builder parity tests).
 * Call with NULL arrays to get the counts first. vertices nv*3, vtx_indices nt*3 (post-build order),
 * arr_bvh n_nodes*10. */
int ref_cpu_mesh(const char* obj_path, int scene_kind, int* nv, int* nt, int* n_nodes, int* n_leaves, int* max_depth, int* max_leaf,
                 float* vertices, int* vtx_indices, float* arr_bvh) {
	Built b;
	build_scene(b, obj_path, scene_kind);
	if (!b.mesh) return -1;
	TriangleMesh* m = b.mesh;
	int nodes = 0, leaves = 0, depth = 0, leaf = 0;
	count_nodes(&m->bvh, 1, nodes, leaves, depth, leaf);
	if (nv) *nv = (int)m->vertices.size();
	if (nt) *nt = (int)m->indices.size();
	if (n_nodes) *n_nodes = nodes;
	if (n_leaves) *n_leaves = leaves;
	if (max_depth) *max_depth = depth;
	if (max_leaf) *max_leaf = leaf;
	if (vertices) for (size_t i = 0; i < m->vertices.size(); i++) for (int k = 0; k < 3; k++) vertices[i * 3 + k] = m->vertices[i][k];
	if (vtx_indices) for (size_t i = 0; i < m->indices.size(); i++) {
		vtx_indices[i * 3 + 0] = m->indices[i].vtxi;
		vtx_indices[i * 3 + 1] = m->indices[i].vtxj;
		vtx_indices[i * 3 + 2] = m->indices[i].vtxk;
	}
	if (arr_bvh) { int next = 1; dump_nodes(&m->bvh, arr_bvh, next, 0); }
	return 0;
}

}
