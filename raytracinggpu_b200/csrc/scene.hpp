/*
 * scene.hpp — C++ host surface over the C ABI, with the reference launcher's vocabulary: Vector, Sphere,
 * TriangleMeshHost (readOBJ / rescale / buildBVH / bvhTreeToArray, optimized.cu:293-535) and Scene
 * (addObject, L, intensity; cpu_launcher.cpp:538-652). Header-only; every method forwards to include/rt_b200.h.
 */
#pragma once
#include "../../include/rt_b200.h"

#include <stdexcept>
#include <string>
#include <vector>

namespace rtb200 {

struct Error : std::runtime_error {
    int code;
    Error(int c) : std::runtime_error(rt_last_error()), code(c) {}
};
inline void check(int rc) {
    if (rc != RT_OK) throw Error(rc);
}

struct Vector {
    float x = 0, y = 0, z = 0;
    Vector() {}
    Vector(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};

/* Sphere(C, R, albedo, mirror, in_refraction_index, out_refraction_index): optimized.cu:119 */
struct Sphere {
    Vector C;
    float R;
    Vector albedo;
    bool mirror;
    float in_refraction_index, out_refraction_index;
    Sphere(const Vector& C_, float R_, const Vector& albedo_, bool mirror_ = false, float n_in = 1.f, float n_out = 1.f)
        : C(C_), R(R_), albedo(albedo_), mirror(mirror_), in_refraction_index(n_in), out_refraction_index(n_out) {}
};

class TriangleMeshHost {
public:
    TriangleMeshHost() { check(rt_mesh_create(&m_)); }
    ~TriangleMeshHost() { rt_mesh_destroy(m_); }
    TriangleMeshHost(const TriangleMeshHost&) = delete;
    TriangleMeshHost& operator=(const TriangleMeshHost&) = delete;

    void readOBJ(const char* obj) { check(rt_mesh_read_obj(m_, obj)); }
    void rescale(float scale, const Vector& offset) {
        const float o[3] = {offset.x, offset.y, offset.z};
        check(rt_mesh_rescale(m_, scale, o));
    }
    /* compute_bbox + buildBVH + bvhTreeToArray in one step (optimized.cu:809-813) */
    void buildBVH() { check(rt_mesh_build_bvh(m_)); }
    /* the same tree built on the device, a level at a time (global_launcher.cu:298-331 builds it in one device thread) */
    double buildBVHDevice(int device = 0) {
        double ms = 0.;
        check(rt_mesh_build_bvh_gpu(m_, device, &ms));
        return ms;
    }
    int n_vertices() const { int32_t a; rt_mesh_counts(m_, &a, nullptr, nullptr); return a; }
    int n_triangles() const { int32_t a; rt_mesh_counts(m_, nullptr, &a, nullptr); return a; }
    int n_bvhs() const { int32_t a; rt_mesh_counts(m_, nullptr, nullptr, &a); return a; }
    const float* vertices() const { return rt_mesh_vertices(m_); }
    const int32_t* indices() const { return rt_mesh_tri_records(m_); }
    const float* arr_bvh() const { return rt_mesh_arr_bvh(m_); }
    rt_mesh* handle() { return m_; }

    Vector albedo = Vector(0.25f, 0.25f, 0.25f); /* optimized.cu:692 */
    bool mirror = false;
    float in_refraction_index = 1.f, out_refraction_index = 1.f;

private:
    rt_mesh* m_ = nullptr;
};

class Scene {
public:
    explicit Scene(int device = 0) { check(rt_scene_create(&s_, device)); }
    ~Scene() { rt_scene_destroy(s_); }
    Scene(const Scene&) = delete;
    Scene& operator=(const Scene&) = delete;

    /* addObject: the object's id is its insertion index (cpu_launcher.cpp:540-543) */
    void addObject(const Sphere& sp) {
        rt_sphere r;
        r.C[0] = sp.C.x; r.C[1] = sp.C.y; r.C[2] = sp.C.z;
        r.R = sp.R;
        r.albedo[0] = sp.albedo.x; r.albedo[1] = sp.albedo.y; r.albedo[2] = sp.albedo.z;
        r.mirror = sp.mirror ? 1 : 0;
        r.n_in = sp.in_refraction_index;
        r.n_out = sp.out_refraction_index;
        r.id = next_id_++;
        spheres_.push_back(r);
        dirty_ = true;
    }
    void addObject(TriangleMeshHost& mesh) {
        const float a[3] = {mesh.albedo.x, mesh.albedo.y, mesh.albedo.z};
        check(rt_scene_set_mesh(s_, mesh.vertices(), mesh.n_vertices(), mesh.indices(), mesh.n_triangles(), mesh.arr_bvh(), mesh.n_bvhs(), a,
                                mesh.mirror ? 1 : 0, mesh.in_refraction_index, mesh.out_refraction_index, next_id_++));
    }
    void setLight(const Vector& L_, float intensity_) {
        L = L_;
        intensity = intensity_;
        dirty_ = true;
    }
    /* push pending addObject / setLight state to the device scene */
    void flush() {
        if (dirty_) {
            check(rt_scene_set_spheres(s_, spheres_.data(), (int32_t)spheres_.size()));
            const float l[3] = {L.x, L.y, L.z};
            check(rt_scene_set_light(s_, l, intensity));
            dirty_ = false;
        }
    }
    /* One frame into a host buffer (H*W*3). */
    rt_stats render(const rt_params& p, uint8_t* rgb, int32_t* hit_obj = nullptr, int32_t* hit_tri = nullptr, float* hit_t = nullptr,
                    uint8_t* shadow = nullptr, uint32_t flags = 0) {
        flush();
        rt_stats st;
        check(rt_render(s_, &p, flags, rgb, hit_obj, hit_tri, hit_t, shadow, &st));
        return st;
    }
    rt_scene* handle() { return s_; }

    Vector L = Vector(-10.f, 20.f, 40.f); /* optimized.cu:681 */
    float intensity = 3e10f;              /* :683 */

private:
    rt_scene* s_ = nullptr;
    std::vector<rt_sphere> spheres_;
    int next_id_ = 0;
    bool dirty_ = true;
};

} // namespace rtb200
