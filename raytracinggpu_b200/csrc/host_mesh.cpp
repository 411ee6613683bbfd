/*
 * host_mesh.cpp — host scene surface: the TriangleMeshHost replacement (rt_mesh).
 *
 * Implements the reference behaviours of
 *   readOBJ          optimized.cu:303-454  (cpu_launcher.cpp:315-493)
 *   rescale          optimized.cu:297-301
 *   compute_bbox     optimized.cu:466-474
 *   buildBVH         optimized.cu:476-510  (cpu_launcher.cpp:190-224)
 *   bvhTreeToArray   optimized.cu:512-534  (array_bvh.cu:733-759)
 * with a different construction: the OBJ file is tokenised from memory, and the BVH is built iteratively
 * (explicit work stack, nodes emitted directly in the flattened pre-order) so that the 10 M-triangle
 * configuration does not depend on the C stack. The results — vertex values, post-build triangle order and
 * the 10-float node array — are identical to the reference's (tests/test_host_mesh.py compares them with
 * oracle/ and with the compiled reference).
 *
 * Build this file with -ffp-contract=off: `v*0.8 + offset` and `(a+b+c)/3` are unfused in the reference.
 */
#include "host_common.h"

#include <algorithm>
#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

struct TriRecord {
    int32_t w[RT_TRI_RECORD_WORDS]; /* vtxi,vtxj,vtxk,uvi,uvj,uvk,ni,nj,nk,group (optimized.cu:140-147) */
};

const float kInf = (float)(1e9 + 9); /* INF as stored into BoundingBox floats, optimized.cu:21,157 */

} // namespace

struct rt_mesh {
    std::vector<rtb::Vec3> vertices;
    std::vector<rtb::Vec3> normals; /* `vn` lines, kept only on request (rt_mesh_keep_normals): realtime_render.cu:489-493 */
    bool keep_normals = false;
    std::vector<TriRecord> tris;
    std::vector<float> arr_bvh; /* n_nodes * 10 */
    int32_t n_nodes = 0, n_leaves = 0, max_depth = 0, max_leaf = 0;
    /* rt_mesh_build_bvh_gpu leaves the post-build arrays on the device (rt_scene_set_mesh_from takes them from there) and fetches the host
     * mirror — the reordered records and arr_bvh — only when somebody asks for it */
    void* device_keep = nullptr;
    bool host_stale = false;
    ~rt_mesh() {
        if (device_keep) rtb::bvh_device_free(device_keep);
    }
};

namespace rtb {
void* mesh_device_handle(rt_mesh* m) { return m ? m->device_keep : nullptr; }
} // namespace rtb

namespace {

/* the host mirror of a device-resident build, on demand */
int materialize(rt_mesh* m) {
    if (!m->host_stale) return RT_OK;
    const int rc = rtb::bvh_device_download(m->device_keep, &m->tris[0].w[0], &m->arr_bvh);
    if (rc != 0) return rtb::fail(RT_ERR_CUDA, "rt_mesh: fetching the device-built BVH failed (CUDA error %d)", rc);
    m->host_stale = false;
    return RT_OK;
}
/* before the mesh changes: the host copy becomes the truth again, the device copy goes */
int take_back(rt_mesh* m) {
    const int rc = materialize(m);
    if (m->device_keep) {
        rtb::bvh_device_free(m->device_keep);
        m->device_keep = nullptr;
    }
    return rc;
}

/* ---- OBJ tokeniser ------------------------------------------------------------------------------------ */

inline bool is_blank(char c) { return c == ' ' || c == '\t' || c == '\r'; }

/* Parse up to `max` floats from [p,end) on one line; returns how many were read. */
int parse_floats(const char* p, const char* end, float* out, int max) {
    int n = 0;
    while (n < max) {
        while (p < end && is_blank(*p)) p++;
        if (p >= end) break;
        char* q = nullptr;
        float v = strtof(p, &q);
        if (q == p) break;
        out[n++] = v;
        p = q;
    }
    return n;
}

/* One face-vertex token `a`, `a/b`, `a/b/c` or `a//c`: the vertex index, and the normal index c (0 = absent) for the
 * loaders that keep it (realtime_render.cu:538-545; optimized.cu drops uv / normal indices). Returns false when no integer
 * starts at p. */
bool parse_face_vertex(const char*& p, const char* end, long& vi, long& ni) {
    while (p < end && is_blank(*p)) p++;
    if (p >= end) return false;
    char* q = nullptr;
    errno = 0;
    long v = strtol(p, &q, 10);
    if (q == p) return false;
    vi = v;
    ni = 0;
    p = q;
    for (int k = 0; k < 2 && p < end && *p == '/'; k++) {
        p++;
        if (p < end && (*p == '-' || (*p >= '0' && *p <= '9'))) {
            const long w = strtol(p, &q, 10);
            if (k == 1) ni = w;
            p = q;
        }
    }
    return true;
}

/* 1-based -> 0-based; negative indices are relative to the vertices read so far (optimized.cu:368). */
inline int32_t resolve(long i, size_t nv) { return i < 0 ? (int32_t)((long)nv + i) : (int32_t)(i - 1); }

int read_obj(rt_mesh* m, const char* path) {
    take_back(m);
    FILE* f = fopen(path, "rb");
    if (!f) return rtb::fail(RT_ERR_IO, "rt_mesh_read_obj: cannot open '%s'", path);
    std::vector<char> buf;
    {
        char chunk[1 << 16];
        size_t n;
        while ((n = fread(chunk, 1, sizeof chunk, f)) > 0) buf.insert(buf.end(), chunk, chunk + n);
    }
    fclose(f);
    buf.push_back('\n');
    m->vertices.clear();
    m->normals.clear();
    m->tris.clear();
    m->arr_bvh.clear();
    m->n_nodes = 0;

    const char* p = buf.data();
    const char* const eof = p + buf.size();
    std::vector<long> poly, poly_n;
    while (p < eof) {
        const char* eol = (const char*)memchr(p, '\n', (size_t)(eof - p));
        if (!eol) eol = eof;
        if (eol - p >= 2 && p[0] == 'v' && p[1] == ' ') {
            float v[6] = {0, 0, 0, 0, 0, 0};
            int n = parse_floats(p + 2, eol, v, 6);
            rtb::Vec3 q{v[0], v[1], v[2]};
            if (n < 6) {
                /* vec*0.8 + (0,-10,0): optimized.cu:342. 6-field (coloured) vertices stay untouched (:332-339). */
                q.x = q.x * 0.8f + 0.f;
                q.y = q.y * 0.8f + -10.f;
                q.z = q.z * 0.8f + 0.f;
            }
            m->vertices.push_back(q);
        } else if (m->keep_normals && eol - p >= 3 && p[0] == 'v' && p[1] == 'n' && p[2] == ' ') {
            float v[3] = {0, 0, 0};
            parse_floats(p + 3, eol, v, 3);
            m->normals.push_back(rtb::Vec3{v[0], v[1], v[2]}); /* realtime_render.cu:489-493: as read */
        } else if (eol - p >= 1 && p[0] == 'f') {
            poly.clear();
            poly_n.clear();
            const char* q = p + 1;
            long vi, ni;
            while (parse_face_vertex(q, eol, vi, ni)) {
                poly.push_back(vi);
                poly_n.push_back(ni);
            }
            const size_t nv = m->vertices.size();
            const size_t nnrm = m->normals.size();
            /* fan triangulation (i0, i_{k-1}, i_k): optimized.cu:398-447 */
            for (size_t k = 2; k < poly.size(); k++) {
                TriRecord t;
                for (int w = 0; w < RT_TRI_RECORD_WORDS; w++) t.w[w] = -1;
                t.w[0] = resolve(poly[0], nv);
                t.w[1] = resolve(poly[k - 1], nv);
                t.w[2] = resolve(poly[k], nv);
                if (m->keep_normals && poly_n[0] && poly_n[k - 1] && poly_n[k]) { /* ni, nj, nk: realtime_render.cu:538-545 */
                    t.w[6] = resolve(poly_n[0], nnrm);
                    t.w[7] = resolve(poly_n[k - 1], nnrm);
                    t.w[8] = resolve(poly_n[k], nnrm);
                }
                m->tris.push_back(t);
            }
        }
        p = eol + 1;
    }
    const int32_t nv = (int32_t)m->vertices.size(), nn = (int32_t)m->normals.size();
    for (const TriRecord& t : m->tris)
        for (int k = 0; k < 3; k++) {
            if (t.w[k] < 0 || t.w[k] >= nv) return rtb::fail(RT_ERR_INVALID, "rt_mesh_read_obj: face index out of range in '%s'", path);
            if (t.w[6 + k] >= nn) return rtb::fail(RT_ERR_INVALID, "rt_mesh_read_obj: normal index out of range in '%s'", path);
        }
    return RT_OK;
}

/* ---- BVH ---------------------------------------------------------------------------------------------- */

struct BuildItem {
    int32_t start, end;
    int32_t parent; /* node index of the parent, -1 for the root */
    int32_t is_right;
    int32_t depth;
};

int build_bvh(rt_mesh* m) {
    const int32_t nt = (int32_t)m->tris.size();
    if (nt >= (1 << 24)) /* indices travel as floats in arr_bvh (optimized.cu:519-529): exact below 2^24 only */
        return rtb::fail(RT_ERR_UNSUPPORTED, "rt_mesh_build_bvh: %d triangles do not fit the float-encoded array BVH (2^24)", nt);
    m->arr_bvh.clear();
    m->n_nodes = m->n_leaves = m->max_depth = m->max_leaf = 0;
    std::vector<float>& arr = m->arr_bvh;
    arr.reserve((size_t)std::max(1, nt / 2) * RT_BVH_NODE_FLOATS);
    const rtb::Vec3* V = m->vertices.data();
    TriRecord* T = m->tris.data();

    std::vector<BuildItem> work;
    work.push_back(BuildItem{0, nt, -1, 0, 1});
    while (!work.empty()) {
        const BuildItem it = work.back();
        work.pop_back();
        /* Popping the left child before the right one emits nodes in the pre-order bvhTreeToArray assigns
         * (optimized.cu:522-533): a node's left child is the next slot, its right child follows the left subtree. */
        const int32_t idx = m->n_nodes++;
        arr.resize((size_t)m->n_nodes * RT_BVH_NODE_FLOATS);
        float* node = &arr[(size_t)idx * RT_BVH_NODE_FLOATS];
        if (it.parent >= 0) arr[(size_t)it.parent * RT_BVH_NODE_FLOATS + (it.is_right ? 1 : 0)] = (float)idx;
        if (it.depth > m->max_depth) m->max_depth = it.depth;

        float mn[3] = {kInf, kInf, kInf}, mx[3] = {-kInf, -kInf, -kInf}; /* compute_bbox :466-474 */
        for (int32_t i = it.start; i < it.end; i++) {
            for (int c = 0; c < 3; c++) {
                const rtb::Vec3& q = V[T[i].w[c]];
                mn[0] = std::min(mn[0], q.x);
                mn[1] = std::min(mn[1], q.y);
                mn[2] = std::min(mn[2], q.z);
                mx[0] = std::max(mx[0], q.x);
                mx[1] = std::max(mx[1], q.y);
                mx[2] = std::max(mx[2], q.z);
            }
        }
        node[0] = -1.f;
        node[1] = -1.f;
        for (int k = 0; k < 3; k++) {
            node[2 + k] = mn[k];
            node[5 + k] = mx[k];
        }
        node[8] = (float)it.start;
        node[9] = (float)it.end;

        const float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
        const int axis = (dx >= dy && dx >= dz) ? 0 : ((dy >= dx && dy >= dz) ? 1 : 2); /* :485-491 */
        const float split = (mn[axis] + mx[axis]) / 2;                                   /* :494 */
        int32_t pivot = it.start;
        for (int32_t i = it.start; i < it.end; i++) {
            const float cen = (V[T[i].w[0]][axis] + V[T[i].w[1]][axis] + V[T[i].w[2]][axis]) / 3; /* :496 */
            if (cen < split) {
                std::swap(T[i], T[pivot]);
                pivot++;
            }
        }
        if (pivot <= it.start || pivot >= it.end - 1 || it.end - it.start < 5) { /* leaf rule :503 */
            m->n_leaves++;
            m->max_leaf = std::max(m->max_leaf, it.end - it.start);
            continue;
        }
        work.push_back(BuildItem{pivot, it.end, idx, 1, it.depth + 1});
        work.push_back(BuildItem{it.start, pivot, idx, 0, it.depth + 1});
    }
    return RT_OK;
}

} // namespace

extern "C" {

int rt_mesh_create(rt_mesh** out) {
    if (!out) return rtb::fail(RT_ERR_INVALID, "rt_mesh_create: out is NULL");
    *out = new (std::nothrow) rt_mesh();
    return *out ? RT_OK : rtb::fail(RT_ERR_NOMEM, "rt_mesh_create: out of memory");
}

void rt_mesh_destroy(rt_mesh* m) { delete m; }

int rt_mesh_read_obj(rt_mesh* m, const char* path) {
    if (!m || !path) return rtb::fail(RT_ERR_INVALID, "rt_mesh_read_obj: NULL argument");
    return read_obj(m, path);
}

int rt_mesh_set_triangles(rt_mesh* m, const float* vertices, int32_t nv, const int32_t* idx, int32_t nt) {
    if (m) take_back(m);
    if (!m || nv < 0 || nt < 0 || (nv > 0 && !vertices) || (nt > 0 && !idx)) return rtb::fail(RT_ERR_INVALID, "rt_mesh_set_triangles: bad argument");
    for (int64_t i = 0; i < (int64_t)nt * 3; i++)
        if (idx[i] < 0 || idx[i] >= nv) return rtb::fail(RT_ERR_INVALID, "rt_mesh_set_triangles: vertex index %d out of range", idx[i]);
    m->vertices.resize(nv);
    for (int32_t i = 0; i < nv; i++) m->vertices[i] = rtb::Vec3{vertices[3 * i], vertices[3 * i + 1], vertices[3 * i + 2]};
    m->tris.resize(nt);
    for (int32_t i = 0; i < nt; i++) {
        for (int w = 0; w < RT_TRI_RECORD_WORDS; w++) m->tris[i].w[w] = -1;
        for (int k = 0; k < 3; k++) m->tris[i].w[k] = idx[3 * i + k];
    }
    m->arr_bvh.clear();
    m->normals.clear();
    m->n_nodes = 0;
    return RT_OK;
}

int rt_mesh_keep_normals(rt_mesh* m, int keep) {
    if (!m) return rtb::fail(RT_ERR_INVALID, "rt_mesh_keep_normals: NULL mesh");
    m->keep_normals = keep != 0;
    return RT_OK;
}

int rt_mesh_set_normals(rt_mesh* m, const float* normals, int32_t nn, const int32_t* normal_indices) {
    if (!m || nn < 0 || (nn > 0 && (!normals || !normal_indices))) return rtb::fail(RT_ERR_INVALID, "rt_mesh_set_normals: bad argument");
    take_back(m);
    if (m->n_nodes > 0) return rtb::fail(RT_ERR_STATE, "rt_mesh_set_normals: attach the normals before the BVH is built (the build reorders the records)");
    const size_t nt = m->tris.size();
    for (size_t i = 0; i < 3 * nt && nn > 0; i++)
        if (normal_indices[i] < 0 || normal_indices[i] >= nn) return rtb::fail(RT_ERR_INVALID, "rt_mesh_set_normals: normal index %d out of range", normal_indices[i]);
    m->normals.resize((size_t)nn);
    for (int32_t i = 0; i < nn; i++) m->normals[i] = rtb::Vec3{normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]};
    for (size_t i = 0; i < nt; i++)
        for (int k = 0; k < 3; k++) m->tris[i].w[6 + k] = nn > 0 ? normal_indices[3 * i + k] : -1;
    return RT_OK;
}

int rt_mesh_normal_count(const rt_mesh* m, int32_t* nn) {
    if (!m || !nn) return rtb::fail(RT_ERR_INVALID, "rt_mesh_normal_count: bad argument");
    *nn = (int32_t)m->normals.size();
    return RT_OK;
}

const float* rt_mesh_normals(const rt_mesh* m) { return (m && !m->normals.empty()) ? &m->normals[0].x : nullptr; }

int rt_mesh_rescale(rt_mesh* m, float scale, const float offset[3]) {
    if (!m || !offset) return rtb::fail(RT_ERR_INVALID, "rt_mesh_rescale: NULL argument");
    take_back(m);
    for (rtb::Vec3& v : m->vertices) { /* vertices[i]*scale + offset, optimized.cu:299 */
        v.x = v.x * scale + offset[0];
        v.y = v.y * scale + offset[1];
        v.z = v.z * scale + offset[2];
    }
    m->arr_bvh.clear();
    m->n_nodes = 0;
    return RT_OK;
}

int rt_mesh_instance(rt_mesh* m, int32_t copies, const float* scales, const float* offsets) {
    if (!m || copies < 1 || !scales || !offsets) return rtb::fail(RT_ERR_INVALID, "rt_mesh_instance: bad argument");
    take_back(m);
    const size_t nv = m->vertices.size(), nt = m->tris.size();
    if ((uint64_t)nt * (uint64_t)copies >= (1u << 24)) return rtb::fail(RT_ERR_UNSUPPORTED, "rt_mesh_instance: %llu triangles exceed 2^24", (unsigned long long)nt * copies);
    std::vector<rtb::Vec3> v(nv * copies);
    std::vector<TriRecord> t(nt * copies);
    for (int32_t c = 0; c < copies; c++) {
        const float s = scales[c], ox = offsets[3 * c], oy = offsets[3 * c + 1], oz = offsets[3 * c + 2];
        for (size_t i = 0; i < nv; i++) {
            const rtb::Vec3& q = m->vertices[i];
            v[c * nv + i] = rtb::Vec3{q.x * s + ox, q.y * s + oy, q.z * s + oz};
        }
        for (size_t i = 0; i < nt; i++) {
            TriRecord r = m->tris[i];
            for (int k = 0; k < 3; k++) r.w[k] += (int32_t)(c * nv);
            t[c * nt + i] = r;
        }
    }
    m->vertices.swap(v);
    m->tris.swap(t);
    m->arr_bvh.clear();
    m->n_nodes = 0;
    return RT_OK;
}

int rt_mesh_build_bvh(rt_mesh* m) {
    if (!m) return rtb::fail(RT_ERR_INVALID, "rt_mesh_build_bvh: NULL mesh");
    take_back(m);
    return build_bvh(m);
}

int rt_mesh_build_bvh_gpu(rt_mesh* m, int device, double* build_ms) {
    if (!m) return rtb::fail(RT_ERR_INVALID, "rt_mesh_build_bvh_gpu: NULL mesh");
    int rc0 = take_back(m);
    if (rc0 != RT_OK) return rc0;
    const int32_t nt = (int32_t)m->tris.size(), nv = (int32_t)m->vertices.size();
    if (build_ms) *build_ms = 0.;
    if (nt < 2) return build_bvh(m); /* nothing to do in parallel */
    if (nt >= (1 << 24)) return rtb::fail(RT_ERR_UNSUPPORTED, "rt_mesh_build_bvh_gpu: %d triangles do not fit the float-encoded array BVH (2^24)", nt);
    std::vector<int32_t> perm;
    std::vector<float> arr;
    int32_t info[4] = {0, 0, 0, 0};
    /* the post-build arrays stay on the device (rt_scene_set_mesh_from takes them from there); the host mirror — records in the
     * builder's order, arr_bvh — is fetched by the accessors when somebody asks for it */
    void* keep = nullptr;
    const int rc = rtb::bvh_build_device(device, &m->vertices[0].x, nv, &m->tris[0].w[0], nt, &perm, &arr, info, build_ms, &keep);
    if (rc != 0) return rtb::fail(RT_ERR_CUDA, "rt_mesh_build_bvh_gpu: CUDA error %d (this library has no CPU fallback for the device builder; rt_mesh_build_bvh is the host builder)", rc);
    m->device_keep = keep;
    m->host_stale = true;
    m->arr_bvh.clear();
    m->n_nodes = info[0];
    m->n_leaves = info[1];
    m->max_depth = info[2];
    m->max_leaf = info[3];
    return RT_OK;
}

int rt_mesh_counts(const rt_mesh* m, int32_t* nv, int32_t* nt, int32_t* n_nodes) {
    if (!m) return rtb::fail(RT_ERR_INVALID, "rt_mesh_counts: NULL mesh");
    if (nv) *nv = (int32_t)m->vertices.size();
    if (nt) *nt = (int32_t)m->tris.size();
    if (n_nodes) *n_nodes = m->n_nodes;
    return RT_OK;
}

const float* rt_mesh_vertices(const rt_mesh* m) { return (m && !m->vertices.empty()) ? &m->vertices[0].x : nullptr; }
const int32_t* rt_mesh_tri_records(const rt_mesh* m) {
    if (m && m->host_stale && materialize(const_cast<rt_mesh*>(m)) != RT_OK) return nullptr;
    return (m && !m->tris.empty()) ? &m->tris[0].w[0] : nullptr;
}
const float* rt_mesh_arr_bvh(const rt_mesh* m) {
    if (m && m->host_stale && materialize(const_cast<rt_mesh*>(m)) != RT_OK) return nullptr;
    return (m && m->n_nodes > 0) ? m->arr_bvh.data() : nullptr;
}

int rt_mesh_bvh_info(const rt_mesh* m, int32_t* n_leaves, int32_t* max_depth, int32_t* max_leaf) {
    if (!m) return rtb::fail(RT_ERR_INVALID, "rt_mesh_bvh_info: NULL mesh");
    if (m->n_nodes == 0) return rtb::fail(RT_ERR_STATE, "rt_mesh_bvh_info: BVH not built");
    if (n_leaves) *n_leaves = m->n_leaves;
    if (max_depth) *max_depth = m->max_depth;
    if (max_leaf) *max_leaf = m->max_leaf;
    return RT_OK;
}

} /* extern "C" */
