// C ABI of libncb200.so (see include/ncb200.h): context, uploads, the fused world update and its stage entry points.
// No torch types, no CPU fallback: every compute entry point needs a CUDA device and fails loudly otherwise.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "ncb_internal.h"

using namespace ncb;

static thread_local std::string g_create_err;

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess) {                                                                        \
            char b__[512];                                                                               \
            snprintf(b__, sizeof b__, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            ctx->err = b__;                                                                              \
            return NCB_ERR_CUDA;                                                                         \
        }                                                                                                \
    } while (0)

#define REQUIRE(cond, code, msg) \
    do {                         \
        if (!(cond)) {           \
            ctx->err = (msg);    \
            return (code);       \
        }                        \
    } while (0)

DevObjects dev_objects(ncb_ctx* c) {
    DevObjects o;
    o.n = c->n;
    o.pos = c->pos.p;
    o.rot = c->rot.p;
    o.type = c->type.p;
    o.param = c->param.p;
    o.groups = c->has_groups ? c->groups.p : nullptr;
    o.qlimit = c->qlimit.p;
    o.ang = c->ang.p;
    o.ang_cs = c->ang_cs.p;
    o.ang_stride = c->ang_stride;
    o.cap_pts = c->has_capsules ? c->cap_pts.p : nullptr;
    return o;
}

int reserve_broad(ncb_ctx* ctx, uint32_t n) {
    CK(ctx->aabb_lo.reserve(n));
    CK(ctx->aabb_hi.reserve(n));
    CK(ctx->keys_a.reserve(n));
    CK(ctx->keys_b.reserve(n));
    CK(ctx->idx_a.reserve(n));
    CK(ctx->idx_b.reserve(n));
    CK(ctx->leaf_lo.reserve(n));
    CK(ctx->leaf_hi.reserve(n));
    CK(ctx->nodes.reserve(4 * (size_t)n));
    CK(ctx->parent.reserve(2 * (size_t)n));
    CK(ctx->flags.reserve(n));
    CK(ctx->cub_tmp.reserve(lbvh_temp_bytes(n) + 256));
    CK(ctx->counters.reserve(1));
    return NCB_OK;
}
int reserve_pairs(ncb_ctx* ctx, size_t cap) {
    CK(ctx->pairs_raw.reserve(cap));
    CK(ctx->pairs.reserve(cap));
    CK(ctx->keys_raw.reserve(cap));
    CK(ctx->pair_algo.reserve(cap));
    CK(ctx->manifold_start.reserve(cap));
    CK(ctx->manifold_count.reserve(cap));
    CK(ctx->epa_queue.reserve(26 * cap));
    CK(ctx->epa_long.reserve(2 * cap));
    CK(ctx->cp_queue.reserve(10 * cap));
    return NCB_OK;
}

int reset_counters(ncb_ctx* ctx) {
    DevCounters z;
    memset(&z, 0, sizeof z);
    for (int k = 0; k < 3; ++k) {
        z.bounds[k] = 0x7f7fffff;          // ordered-int of +FLT_MAX
        z.bounds[3 + k] = (int)0x80800000;  // ordered-int of -FLT_MAX  (0xff7fffff ^ 0x7fffffff)
    }
    *ctx->h_counters = z;
    CK(cudaMemcpyAsync(ctx->counters.p, ctx->h_counters, sizeof z, cudaMemcpyHostToDevice, ctx->stream));
    return NCB_OK;
}

extern "C" {

const char* ncb_version(void) { return "ncb200 0.1 (sm_100a)"; }

int ncb_create(int device, ncb_ctx** out) {
    if (!out) return NCB_ERR_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_err = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (ncb200 has no CPU fallback)";
        return NCB_ERR_CUDA;
    }
    if (device < 0 || device >= count) {
        g_create_err = "device index out of range";
        return NCB_ERR_ARG;
    }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        g_create_err = cudaGetErrorString(e);
        return NCB_ERR_CUDA;
    }
    ncb_ctx* c = new ncb_ctx;
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&c->h_counters, sizeof(DevCounters));
    if (e == cudaSuccess) e = cudaMallocHost((void**)&c->h_snap, sizeof(DevCounters));
    if (e == cudaSuccess) e = cudaMallocHost((void**)&c->h_snap_n, (1 + NCB_MAN_PARTS) * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_pairs, cudaEventDisableTiming);
    for (int k = 0; k <= NCB_MAN_PARTS && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&c->ev_snap[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming);
    if (e == cudaSuccess) e = c->snap.reserve(1 + NCB_MAN_PARTS);
    if (e == cudaSuccess && !getenv("NCB_NO_SIDE_STREAM")) {
        e = cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_tier1, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_tier2, cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        g_create_err = cudaGetErrorString(e);
        delete c;
        return NCB_ERR_CUDA;
    }
    c->stream = c->own_stream;
    *out = c;
    return NCB_OK;
}

static void free_hulls(ncb_ctx* c) {
    for (void* p : c->hull_allocs) cudaFree(p);
    c->hull_allocs.clear();
    memset(&c->hulls, 0, sizeof c->hulls);
}

static void p2p_close(ncb_ctx* ctx);
void ncb_destroy(ncb_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    p2p_close(c);  // unmap the peers' buffers (the own ones go with the context)
    free_hulls(c);
    c->ang_cs.release();
    c->pos.release(), c->qlimit.release(), c->ang.release(), c->rot.release(), c->param.release(), c->type.release(), c->groups.release();
    c->aabb_lo.release(), c->aabb_hi.release(), c->keys_a.release(), c->keys_b.release(), c->idx_a.release(), c->idx_b.release();
    c->cub_tmp.release(), c->leaf_lo.release(), c->leaf_hi.release(), c->nodes.release(), c->parent.release(), c->flags.release();
    c->pairs_raw.release(), c->pairs.release(), c->keys_raw.release(), c->pair_algo.release(), c->counters.release();
    c->contacts.release(), c->manifold_start.release(), c->manifold_count.release(), c->pair_index.release();
    c->epa_queue.release(), c->cp_queue.release();
    c->qkind.release(), c->prox.release();
    if (c->timer.created)
        for (int i = 0; i <= StageTimer::MAX; ++i) cudaEventDestroy(c->timer.ev[i]);
    if (c->h_counters) cudaFreeHost(c->h_counters);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->ev_pairs) cudaEventDestroy(c->ev_pairs);
    for (int k = 0; k <= NCB_MAN_PARTS; ++k)
        if (c->ev_snap[k]) cudaEventDestroy(c->ev_snap[k]);
    if (c->h_snap_n) cudaFreeHost(c->h_snap_n);
    if (c->ev_copy) cudaEventDestroy(c->ev_copy);
    if (c->h_snap) cudaFreeHost(c->h_snap);
    c->snap.release();
    if (c->side_stream) cudaStreamDestroy(c->side_stream);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->ev_tier1) cudaEventDestroy(c->ev_tier1);
    if (c->ev_tier2) cudaEventDestroy(c->ev_tier2);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

const char* ncb_last_error(const ncb_ctx* c) { return c ? c->err.c_str() : g_create_err.c_str(); }

int ncb_set_stream(ncb_ctx* ctx, void* s) {
    if (!ctx) return NCB_ERR_ARG;
    cudaStream_t next = s ? (cudaStream_t)s : ctx->own_stream;
    if (next != ctx->stream) {
        // work already enqueued on the outgoing stream (uploads of ncb_set_objects / ncb_set_hulls, their fill kernels) must be
        // complete before anything on the new stream reads it: the two streams are not ordered with each other
        CK(cudaSetDevice(ctx->device));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    ctx->stream = next;
    return NCB_OK;
}
void* ncb_get_stream(ncb_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int ncb_synchronize(ncb_ctx* ctx) {
    if (!ctx) return NCB_ERR_ARG;
    CK(cudaStreamSynchronize(ctx->stream));
    return NCB_OK;
}

int ncb_profile_enable(ncb_ctx* ctx, int on) {
    if (!ctx) return NCB_ERR_ARG;
    ctx->timer.enabled = on != 0;
    return NCB_OK;
}
int ncb_profile_get(ncb_ctx* ctx, const char** names, float* ms, uint32_t* launches) {
    if (!ctx) return NCB_ERR_ARG;
    StageTimer& t = ctx->timer;
    if (!t.enabled || t.n == 0) return 0;
    cudaEventSynchronize(t.ev[t.n]);
    for (int i = 0; i < t.n; ++i) {
        float v = 0;
        cudaEventElapsedTime(&v, t.ev[i], t.ev[i + 1]);
        if (names) names[i] = t.names[i];
        if (ms) ms[i] = v;
        if (launches) launches[i] = t.launches[i];
    }
    return t.n;
}

}  // extern "C"

// ---- uploads -----------------------------------------------------------------------------------------------
template <typename T>
static cudaError_t upload(ncb_ctx* c, const T* host, size_t count, const T** dev_out) {
    T* d = nullptr;
    size_t bytes = (count ? count : 1) * sizeof(T);
    cudaError_t e = cudaMalloc((void**)&d, bytes);
    if (e != cudaSuccess) return e;
    c->hull_allocs.push_back(d);
    if (count) e = cudaMemcpyAsync(d, host, count * sizeof(T), cudaMemcpyHostToDevice, c->stream);
    *dev_out = d;
    return e;
}

extern "C" {

int ncb_set_hulls(ncb_ctx* ctx, const ncb_hull_library* L) {
    if (!ctx || !L) return NCB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    free_hulls(ctx);
    uint32_t nh = L->n_hulls;
    if (nh == 0) return NCB_OK;
    uint32_t nv = L->vert_off[nh], nf = L->face_off[nh], ne = L->edge_off[nh], nfa = L->fadj_off[nh], nva = L->vadj_off[nh];
    for (uint32_t h = 0; h < nh; ++h) {
        REQUIRE(L->vert_off[h + 1] - L->vert_off[h] >= 1 && L->vert_off[h + 1] - L->vert_off[h] <= 64, NCB_ERR_UNSUPPORTED,
                "convex hulls must have 1..64 vertices");
        REQUIRE(L->face_off[h + 1] > L->face_off[h], NCB_ERR_UNSUPPORTED, "convex hull without faces");
    }
    for (uint32_t f = 0; f < nf; ++f) REQUIRE(L->face_num[f] <= 16, NCB_ERR_UNSUPPORTED, "convex hull faces must have <= 16 vertices");
    DevHulls& H = ctx->hulls;
    H.n_hulls = nh;
    CK(upload(ctx, L->vert_off, nh + 1, &H.vert_off));
    CK(upload(ctx, L->face_off, nh + 1, &H.face_off));
    CK(upload(ctx, L->edge_off, nh + 1, &H.edge_off));
    CK(upload(ctx, L->fadj_off, nh + 1, &H.fadj_off));
    CK(upload(ctx, L->vadj_off, nh + 1, &H.vadj_off));
    CK(upload(ctx, L->points, 3 * (size_t)nv, &H.points));
    {
        std::vector<float4> padded(nv);
        for (uint32_t v = 0; v < nv; ++v) padded[v] = make_float4(L->points[3 * (size_t)v], L->points[3 * (size_t)v + 1], L->points[3 * (size_t)v + 2], 0.f);
        CK(upload(ctx, padded.data(), nv, &H.points4));
        CK(cudaStreamSynchronize(ctx->stream));  // `padded` goes away
    }
    CK(upload(ctx, L->vert_first_adj, nv, &H.vert_first_adj));
    CK(upload(ctx, L->vert_num_adj, nv, &H.vert_num_adj));
    CK(upload(ctx, L->face_first, nf, &H.face_first));
    CK(upload(ctx, L->face_num, nf, &H.face_num));
    CK(upload(ctx, L->face_normal, 3 * (size_t)nf, &H.face_normal));
    CK(upload(ctx, L->vertices_adj_to_face, nfa, &H.vaf));
    CK(upload(ctx, L->edges_adj_to_face, nfa, &H.eaf));
    CK(upload(ctx, L->edge_vertices, 2 * (size_t)ne, &H.edge_vertices));
    CK(upload(ctx, L->edge_faces, 2 * (size_t)ne, &H.edge_faces));
    CK(upload(ctx, L->edge_dir, 3 * (size_t)ne, &H.edge_dir));
    CK(upload(ctx, L->faces_adj_to_vertex, nva, &H.fav));
    CK(upload(ctx, L->edges_adj_to_vertex, nva, &H.eav));
    CK(cudaStreamSynchronize(ctx->stream));
    return NCB_OK;
}

int ncb_set_objects(ncb_ctx* ctx, const ncb_objects* o) {
    if (!ctx || !o) return NCB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    uint32_t n = o->n;
    REQUIRE(n == 0 || (o->pos && o->rot && o->shape_type && o->shape_param && o->query_limit && o->ang_pred), NCB_ERR_ARG,
            "ncb_set_objects: null array");
    cudaStream_t s = ctx->stream;
    // validate before any device state is touched: only the shapes of the path exist on the device (an unknown id would be
    // dispatched as something else), and a hull object must name a hull of the uploaded library
    bool has_capsules = false;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t t = o->shape_type[i];
        REQUIRE(t <= NCB_SHAPE_CAPSULE, NCB_ERR_UNSUPPORTED, "ncb_set_objects: shape_type must be NCB_SHAPE_BALL / CUBOID / CONVEX_HULL / PLANE / CAPSULE");
        has_capsules = has_capsules || t == NCB_SHAPE_CAPSULE;
        if (t == NCB_SHAPE_CONVEX_HULL) {
            float h = o->shape_param[4 * (size_t)i];
            REQUIRE(h >= 0.f && h < (float)ctx->hulls.n_hulls && h == (float)(uint32_t)h, NCB_ERR_ARG,
                    "ncb_set_objects: convex-hull object names a hull id that is not in the uploaded library (ncb_set_hulls)");
        }
    }
    CK(ctx->pos.reserve(3 * (size_t)n));
    CK(ctx->rot.reserve(n));
    CK(ctx->type.reserve(n));
    CK(ctx->param.reserve(n));
    CK(ctx->qlimit.reserve(n));
    CK(ctx->ang.reserve(n));
    CK(ctx->ang_cs.reserve(n));
    // the large copies go first; the host work below (angular prediction table) overlaps with them
    if (n) {
        CK(cudaMemcpyAsync(ctx->pos.p, o->pos, 12 * (size_t)n, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ctx->rot.p, o->rot, 16 * (size_t)n, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ctx->type.p, o->shape_type, 4 * (size_t)n, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ctx->param.p, o->shape_param, 16 * (size_t)n, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ctx->qlimit.p, o->query_limit, 4 * (size_t)n, cudaMemcpyHostToDevice, s));
    }
    if (has_capsules) {  // the capsule segments as 2-point hulls [b, a] for GJK / EPA (capsule.cuh)
        CK(ctx->cap_pts.reserve(6 * (size_t)n));
        CK(launch_fill_cap_pts(ctx, n));
    }
    ctx->has_capsules = has_capsules;
    ctx->has_groups = o->groups != nullptr;
    if (o->groups && n) {
        CK(ctx->groups.reserve(3 * (size_t)n));
        CK(cudaMemcpyAsync(ctx->groups.p, o->groups, 12 * (size_t)n, cudaMemcpyHostToDevice, s));
    }
    // cos/sin of the angular prediction: the reference evaluates them with libm (f32::cos / f32::sin in
    // Cuboid::support_feature_toward, cuboid.rs:317,331 and ConvexHull::support_feature_id_toward_eps, convex.rs:389);
    // CUDA's cosf/sinf are not the same function, so they are evaluated here, once per distinct value.  Only this table
    // goes to the device (the raw angles are not read by any kernel).
    bool uniform_ang = true;
    for (uint32_t i = 1; i < n && uniform_ang; ++i) uniform_ang = o->ang_pred[i] == o->ang_pred[0];
    ctx->ang_stride = uniform_ang ? 0u : 1u;
    ctx->h_ang_cs.resize(uniform_ang ? (n ? 1 : 0) : n);
    if (uniform_ang) {
        if (n) ctx->h_ang_cs[0] = make_float2(cosf(o->ang_pred[0]), sinf(o->ang_pred[0]));
    } else {
        float last = 0.f;
        float2 cs = make_float2(1.f, 0.f);
        for (uint32_t i = 0; i < n; ++i) {
            float a = o->ang_pred[i];
            if (a != last) {
                last = a;
                cs = make_float2(cosf(a), sinf(a));
            }
            ctx->h_ang_cs[i] = cs;
        }
    }
    if (n) CK(cudaMemcpyAsync(ctx->ang_cs.p, ctx->h_ang_cs.data(), 8 * ctx->h_ang_cs.size(), cudaMemcpyHostToDevice, s));
    ctx->has_prox = false;  // query types belong to an object set: every object is Contacts after ncb_set_objects (ncb200.h)
    ctx->n = n;
    CK(reserve_broad(ctx, n) == NCB_OK ? cudaSuccess : cudaErrorMemoryAllocation);
    return NCB_OK;
}

int ncb_set_query_types(ncb_ctx* ctx, uint32_t n, const uint8_t* kinds) {
    if (!ctx) return NCB_ERR_ARG;
    REQUIRE(n == ctx->n, NCB_ERR_ARG, "ncb_set_query_types: n differs from the object count");
    CK(cudaSetDevice(ctx->device));
    bool any = false;
    for (uint32_t i = 0; kinds && i < n; ++i) {
        REQUIRE(kinds[i] <= 1, NCB_ERR_ARG, "ncb_set_query_types: kind must be 0 (Contacts) or 1 (Proximity)");
        any = any || kinds[i] != 0;
    }
    REQUIRE(!(any && ctx->has_capsules), NCB_ERR_UNSUPPORTED, "ncb_set_query_types: sensors in a world with capsules are not supported on the device");
    ctx->has_prox = any;
    if (any) {
        CK(ctx->qkind.reserve(n));
        CK(cudaMemcpyAsync(ctx->qkind.p, kinds, n, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));  // the caller's array may go away
    }
    return NCB_OK;
}

int ncb_set_positions(ncb_ctx* ctx, uint32_t n, const float* pos, const float* rot) {
    if (!ctx) return NCB_ERR_ARG;
    REQUIRE(n == ctx->n, NCB_ERR_ARG, "ncb_set_positions: n differs from the object count");
    CK(cudaSetDevice(ctx->device));
    if (n) {
        CK(cudaMemcpyAsync(ctx->pos.p, pos, 12 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->rot.p, rot, 16 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    }
    return NCB_OK;
}

int ncb_set_positions_range(ncb_ctx* ctx, uint32_t begin, uint32_t count, const float* pos, const float* rot) {
    if (!ctx) return NCB_ERR_ARG;
    REQUIRE((uint64_t)begin + count <= ctx->n, NCB_ERR_ARG, "ncb_set_positions_range: range exceeds the object count");
    CK(cudaSetDevice(ctx->device));
    if (count) {
        CK(cudaMemcpyAsync(ctx->pos.p + 3 * (size_t)begin, pos, 12 * (size_t)count, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->rot.p + begin, rot, 16 * (size_t)count, cudaMemcpyHostToDevice, ctx->stream));
    }
    return NCB_OK;
}

// ---- capsule segments ----------------------------------------------------------------------------------------
__global__ void k_fill_cap_pts(const uint32_t* __restrict__ type, const float4* __restrict__ param, uint32_t n, float* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float hh = (type[i] & NCB_TYPE_MASK) == NCB_SHAPE_CAPSULE ? param[i].x : 0.f;
    float* d = out + 6 * (size_t)i;  // b = (0, hh, 0) then a = (0, -hh, 0): Segment::local_support_point as a first-maximum scan
    d[0] = 0.f, d[1] = hh, d[2] = 0.f, d[3] = 0.f, d[4] = -hh, d[5] = 0.f;
}
extern "C++" cudaError_t launch_fill_cap_pts(ncb_ctx* ctx, uint32_t n) {
    if (n == 0) return cudaSuccess;
    k_fill_cap_pts<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->type.p, ctx->param.p, n, ctx->cap_pts.p);
    return cudaGetLastError();
}

// ---- stage entry points ------------------------------------------------------------------------------------
__global__ void k_unpack_aabb(const float4* lo, const float4* hi, uint32_t n, float* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 a = lo[i], b = hi[i];
    float* d = out + 6 * (size_t)i;
    d[0] = a.x, d[1] = a.y, d[2] = a.z, d[3] = b.x, d[4] = b.y, d[5] = b.z;
}
__global__ void k_pack_aabb(const float* in, uint32_t n, float4* lo, float4* hi) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* s = in + 6 * (size_t)i;
    lo[i] = make_float4(s[0], s[1], s[2], 0.f);
    hi[i] = make_float4(s[3], s[4], s[5], __uint_as_float(0u));
}

int ncb_compute_aabbs(ncb_ctx* ctx, float margin, int mode, float* out_minmax) {
    if (!ctx || !out_minmax) return NCB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    uint32_t n = ctx->n;
    if (n == 0) return NCB_OK;
    CK(launch_aabbs(ctx, dev_objects(ctx), margin, mode, 0, n));
    float* tmp = reinterpret_cast<float*>(ctx->nodes.p);  // 16 n floats available
    k_unpack_aabb<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->aabb_lo.p, ctx->aabb_hi.p, n, tmp);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_minmax, tmp, 24 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return NCB_OK;
}

extern "C++" uint32_t* trav_overflow_counter(ncb_ctx* ctx) {
    if (!ctx->trav_overflow.p) {
        if (ctx->trav_overflow.reserve(4) != cudaSuccess) return nullptr;
        cudaMemsetAsync(ctx->trav_overflow.p, 0, 16, ctx->stream);
    }
    return ctx->trav_overflow.p;
}

int ncb_set_kinematics(ncb_ctx* ctx, int on) {
    if (!ctx) return NCB_ERR_ARG;
    ctx->want_kinematics = on != 0;
    if (!on) ctx->have_kinematics = false;
    return NCB_OK;
}
int ncb_world_fetch_kinematics(ncb_ctx* ctx, ncb_kinematic* out, uint32_t cap_contacts) {
    if (!ctx || (cap_contacts && !out)) return NCB_ERR_ARG;
    REQUIRE(ctx->have_kinematics, NCB_ERR_STATE, "ncb_world_fetch_kinematics: call ncb_set_kinematics(ctx, 1) before the update");
    CK(cudaSetDevice(ctx->device));
    uint32_t nc = ctx->last_n_contacts, w = nc < cap_contacts ? nc : cap_contacts;
    if (w) {
        CK(cudaMemcpyAsync(out, ctx->kinematics.p, sizeof(ncb_kinematic) * (size_t)w, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return nc > cap_contacts ? 1 : NCB_OK;
}

int ncb_traversal_overflows(ncb_ctx* ctx, uint32_t* out) {
    if (!ctx || !out) return NCB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    *out = ctx->last_counters.stack_overflow;  // pair search of the last update / ncb_broad_phase
    if (!ctx->trav_overflow.p) return NCB_OK;
    uint32_t q = 0;
    CK(cudaMemcpyAsync(&q, ctx->trav_overflow.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *out += q;
    return NCB_OK;
}

extern "C++" int read_counters(ncb_ctx* ctx) {
    CK(cudaMemcpyAsync(ctx->h_counters, ctx->counters.p, sizeof(DevCounters), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->last_counters = *ctx->h_counters;
    return NCB_OK;
}

int ncb_broad_phase(ncb_ctx* ctx, uint32_t n, const float* aabb_minmax, const uint32_t* groups, uint32_t* out_pairs, uint32_t cap_pairs,
                    uint32_t* n_pairs) {
    if (!ctx || !n_pairs || (n && !aabb_minmax)) return NCB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    *n_pairs = 0;
    if (n == 0) return NCB_OK;
    int r = reserve_broad(ctx, n);
    if (r) return r;
    cudaStream_t s = ctx->stream;
    float* tmp = reinterpret_cast<float*>(ctx->nodes.p);
    CK(cudaMemcpyAsync(tmp, aabb_minmax, 24 * (size_t)n, cudaMemcpyHostToDevice, s));
    k_pack_aabb<<<(n + 255) / 256, 256, 0, s>>>(tmp, n, ctx->aabb_lo.p, ctx->aabb_hi.p);
    CK(cudaGetLastError());
    const uint32_t* dgroups = nullptr;
    if (groups) {
        CK(ctx->pair_index.reserve(3 * (size_t)n));
        CK(cudaMemcpyAsync(ctx->pair_index.p, groups, 12 * (size_t)n, cudaMemcpyHostToDevice, s));
        dgroups = ctx->pair_index.p;
    }
    size_t cap = cap_pairs ? cap_pairs : 1;
    r = reserve_pairs(ctx, cap);
    if (r) return r;
    r = reset_counters(ctx);
    if (r) return r;
    CK(launch_lbvh_build(ctx, n, nullptr));
    CK(launch_pair_search(ctx, n, dgroups, 0, n, (uint32_t)cap));
    r = read_counters(ctx);
    if (r) return r;
    uint32_t found = ctx->last_counters.n_pairs;
    *n_pairs = found;
    uint32_t w = found < cap_pairs ? found : cap_pairs;
    if (out_pairs && w) {
        CK(cudaMemcpyAsync(out_pairs, ctx->pairs_raw.p, 8 * (size_t)w, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    }
    return found > cap_pairs ? 1 : NCB_OK;
}

static void fill_counts(ncb_ctx* ctx, ncb_update_counts* counts) {
    if (!counts) return;
    const DevCounters& c = ctx->last_counters;
    memset(counts, 0, sizeof *counts);
    counts->n_pairs = c.n_pairs;
    counts->n_contacts = c.n_contacts;
    counts->n_contact_pairs = c.n_contact_pairs;
    counts->epa_overflow = c.epa_overflow;
    counts->ref_panics = c.ref_panics;
    counts->n_algo[NCB_ALGO_BALL_BALL] = c.key_hist[K_BALL_BALL];
    counts->n_algo[NCB_ALGO_PLANE_BALL] = c.key_hist[K_PLANE_BALL];
    counts->n_algo[NCB_ALGO_PLANE_CONVEX] = c.key_hist[K_PLANE_CUBOID] + c.key_hist[K_PLANE_HULL];
    counts->n_algo[NCB_ALGO_BALL_CONVEX] = c.key_hist[K_BALL_CUBOID] + c.key_hist[K_BALL_HULL];
    counts->n_algo[NCB_ALGO_CONVEX_CONVEX] = c.key_hist[K_CUBOID_CUBOID] + c.key_hist[K_CUBOID_HULL] + c.key_hist[K_HULL_HULL];
    counts->n_algo[NCB_ALGO_NONE] = c.key_hist[K_NONE];
    counts->n_epa_pairs = c.epa_cursor[K_CUBOID_CUBOID] - c.key_start[K_CUBOID_CUBOID];
    counts->n_manifold_jobs = c.cp_cursor[K_CUBOID_CUBOID] - c.key_start[K_CUBOID_CUBOID];
    counts->n_proximity_pairs = c.key_hist[K_PROX_BALL_BALL] + c.key_hist[K_PROX_PLANE] + c.key_hist[K_PROX_SM] + c.key_hist[K_PROX_SM_HULL];
    for (int k = 0; k < 3; ++k) counts->n_proximity[k] = c.prox_hist[k];
    counts->n_capsule_pairs[0] = c.key_hist[K_CAPSULE_CAPSULE];
    counts->n_capsule_pairs[1] = c.key_hist[K_CAPSULE_BALL] + c.key_hist[K_CAPSULE_PLANE] + c.key_hist[K_CAPSULE_CUBOID] + c.key_hist[K_CAPSULE_HULL];
    counts->stack_overflow = c.stack_overflow;
    counts->n_epa_restarts = c.epa_long_n;
}

int ncb_proximity(ncb_ctx* ctx, uint32_t n_pairs, const uint32_t* pairs, const float* margins, uint8_t* out) {
    if (!ctx || (n_pairs && (!pairs || !out))) return NCB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (n_pairs == 0) return NCB_OK;
    REQUIRE(ctx->n > 0, NCB_ERR_STATE, "ncb_proximity: call ncb_set_objects first");
    for (uint32_t p = 0; p < 2 * n_pairs; ++p) REQUIRE(pairs[p] < ctx->n, NCB_ERR_ARG, "ncb_proximity: object index out of range");
    cudaStream_t s = ctx->stream;
    CK(ctx->pairs_raw.reserve(n_pairs));
    CK(ctx->prox.reserve(n_pairs));
    CK(cudaMemcpyAsync(ctx->pairs_raw.p, pairs, 8 * (size_t)n_pairs, cudaMemcpyHostToDevice, s));
    const float* d_margins = nullptr;
    if (margins) {
        // the margins ride in the manifold_start buffer (same element size), unused by this call
        CK(ctx->manifold_start.reserve(n_pairs));
        CK(cudaMemcpyAsync(ctx->manifold_start.p, margins, 4 * (size_t)n_pairs, cudaMemcpyHostToDevice, s));
        d_margins = reinterpret_cast<const float*>(ctx->manifold_start.p);
    }
    CK(launch_proximity_batch(ctx, dev_objects(ctx), ctx->pairs_raw.p, n_pairs, d_margins, ctx->prox.p));
    CK(cudaMemcpyAsync(out, ctx->prox.p, n_pairs, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return NCB_OK;
}

int ncb_generate_contacts(ncb_ctx* ctx, uint32_t n_pairs, const uint32_t* pairs, ncb_contact* out_contacts, uint32_t cap_contacts,
                          uint32_t* n_contacts, uint32_t* manifold_start, uint8_t* manifold_count, uint8_t* algo) {
    if (!ctx || !n_contacts || (n_pairs && !pairs)) return NCB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    *n_contacts = 0;
    if (n_pairs == 0) return NCB_OK;
    REQUIRE(ctx->n > 0, NCB_ERR_STATE, "ncb_generate_contacts: call ncb_set_objects first");
    for (uint32_t p = 0; p < 2 * n_pairs; ++p) REQUIRE(pairs[p] < ctx->n, NCB_ERR_ARG, "ncb_generate_contacts: object index out of range");
    cudaStream_t s = ctx->stream;
    int r = reserve_pairs(ctx, n_pairs);
    if (r) return r;
    CK(ctx->pair_index.reserve(n_pairs));
    CK(ctx->counters.reserve(1));
    size_t capc = cap_contacts ? cap_contacts : 1;
    CK(ctx->contacts.reserve(capc));
    if (ctx->want_kinematics) CK(ctx->kinematics.reserve(capc));
    ctx->have_kinematics = false;
    r = reset_counters(ctx);
    if (r) return r;
    CK(cudaMemcpyAsync(ctx->pairs_raw.p, pairs, 8 * (size_t)n_pairs, cudaMemcpyHostToDevice, s));
    CK(launch_classify_pairs(ctx, ctx->pairs_raw.p, n_pairs));
    CK(launch_pair_sort(ctx, n_pairs, ctx->pair_index.p));
    CK(launch_narrow_phase(ctx, dev_objects(ctx), ctx->pairs.p, ctx->pair_index.p, n_pairs, (uint32_t)capc));
    r = read_counters(ctx);
    if (r) return r;
    uint32_t nc = ctx->last_counters.n_contacts;
    *n_contacts = nc;
    ctx->last_n_contacts = nc < cap_contacts ? nc : cap_contacts;
    ctx->have_kinematics = ctx->want_kinematics;
    uint32_t w = nc < cap_contacts ? nc : cap_contacts;
    if (out_contacts && w) CK(cudaMemcpyAsync(out_contacts, ctx->contacts.p, sizeof(ncb_contact) * (size_t)w, cudaMemcpyDeviceToHost, s));
    if (manifold_start) CK(cudaMemcpyAsync(manifold_start, ctx->manifold_start.p, 4 * (size_t)n_pairs, cudaMemcpyDeviceToHost, s));
    if (manifold_count) CK(cudaMemcpyAsync(manifold_count, ctx->manifold_count.p, (size_t)n_pairs, cudaMemcpyDeviceToHost, s));
    if (algo) {
        // algo per ORIGINAL pair: scatter back on the host from the sorted order
        std::vector<uint8_t> sorted_algo(n_pairs);
        std::vector<uint32_t> index(n_pairs);
        CK(cudaMemcpyAsync(sorted_algo.data(), ctx->pair_algo.p, n_pairs, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(index.data(), ctx->pair_index.p, 4 * (size_t)n_pairs, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        for (uint32_t p = 0; p < n_pairs; ++p) algo[index[p]] = sorted_algo[p];
    }
    CK(cudaStreamSynchronize(s));
    return nc > cap_contacts ? 1 : NCB_OK;
}

// ---- fused hot path ----------------------------------------------------------------------------------------
// n_local / handle_map / my_rank: the spatially sharded variant runs on the selected (owned + ghost) boxes that the caller
// put in ctx->aabb_lo / aabb_hi, reports global handles through handle_map and filters pairs by ownership.
static int update_after_aabbs(ncb_ctx* ctx, uint32_t q_begin, uint32_t q_end, uint32_t n_local = 0xffffffffu, const uint32_t* handle_map = nullptr,
                              int my_rank = -1) {
    uint32_t n = n_local == 0xffffffffu ? ctx->n : n_local;
    // capacities: grow-on-overflow, remembered across calls
    size_t cap_pairs = ctx->cap_pairs_hint ? ctx->cap_pairs_hint : (size_t)6 * n + 1024;
    size_t cap_contacts = ctx->cap_contacts_hint ? ctx->cap_contacts_hint : (size_t)6 * n + 1024;
    for (int attempt = 0; attempt < 3; ++attempt) {
        int r = reserve_pairs(ctx, cap_pairs);
        if (r) return r;
        CK(ctx->contacts.reserve(cap_contacts));
        if (ctx->want_kinematics) CK(ctx->kinematics.reserve(cap_contacts));
        ctx->have_kinematics = false;
        r = reset_counters(ctx);
        if (r) return r;
        if (!ctx->timer_external || attempt > 0) timer_begin(ctx);
        CK(launch_lbvh_build(ctx, n, handle_map));
        CK(launch_pair_search(ctx, n, ctx->has_groups ? ctx->groups.p : nullptr, q_begin, q_end, (uint32_t)cap_pairs, my_rank));
        timer_mark(ctx, "pair_search", 2);
        if (ctx->has_prox) {  // sensors: their pairs move to the proximity keys; statuses default to NONE
            CK(ctx->prox.reserve(cap_pairs));
            CK(cudaMemsetAsync(ctx->prox.p, 0xff, cap_pairs, ctx->stream));
            CK(launch_prox_rekey(ctx, (uint32_t)cap_pairs));
        }
        // sharded updates read their narrow-phase operands from compact rank-local arrays (broad.cu: k_gather_local_objects)
        static const bool local_ops_ok = getenv("NCB_SHARD_GLOBAL_OPERANDS") == nullptr;
        const bool local_ops = local_ops_ok && handle_map != nullptr && n < ctx->n;
        DevObjects objs = dev_objects(ctx);
        if (local_ops) {
            CK(ctx->pairs_local.reserve(cap_pairs));
            CK(launch_gather_local_objects(ctx, handle_map, n, dev_objects(ctx), &objs));
        }
        CK(launch_pair_sort(ctx, (uint32_t)cap_pairs, nullptr, local_ops ? ctx->local_of.p : nullptr, local_ops ? ctx->pairs_local.p : nullptr));
        timer_mark(ctx, "pair_sort", local_ops ? 4 : 3);
        if (ctx->early.active) CK(cudaEventRecord(ctx->ev_pairs, ctx->stream));
        CK(launch_narrow_phase(ctx, objs, local_ops ? ctx->pairs_local.p : ctx->pairs.p, nullptr, (uint32_t)cap_pairs, (uint32_t)cap_contacts));
        if (ctx->early.active) {
            // Everything of this update is enqueued.  While the narrow phase runs, the copy stream ships what is already
            // final: the sorted pair list (+ algorithm per pair) once the pair sort is done, the contacts written by the
            // kernels that finished before the convex-convex EPA / manifold phases (snapshot 0), then the contacts of each
            // part of the manifold kernel while the next part computes (snapshots 1 .. NCB_MAN_PARTS); ncb_world_fetch is
            // left with the per-pair manifold index.  Contact slots are allocated with one atomic counter, so what a
            // snapshot covers is a contiguous, final prefix of the contact array.
            ncb_ctx::EarlyFetch& ef = ctx->early;
            ef.pairs_done = ef.contacts_done = 0;
            cudaStream_t cs = ctx->copy_stream;
            CK(cudaStreamWaitEvent(cs, ctx->ev_pairs, 0));
            CK(cudaMemcpyAsync(ctx->h_snap, ctx->counters.p, sizeof(DevCounters), cudaMemcpyDeviceToHost, cs));
            CK(cudaStreamSynchronize(cs));
            uint32_t np = ctx->h_snap->n_pairs;
            if (np <= cap_pairs) {
                uint32_t wp = np < ef.cap_pairs ? np : ef.cap_pairs;
                if (ef.pairs && wp) CK(cudaMemcpyAsync(ef.pairs, ctx->pairs.p, 8 * (size_t)wp, cudaMemcpyDeviceToHost, cs));
                if (ef.algo && wp) CK(cudaMemcpyAsync(ef.algo, ctx->pair_algo.p, wp, cudaMemcpyDeviceToHost, cs));
                ef.pairs_done = wp;
                for (int k = 0; k <= NCB_MAN_PARTS; ++k) {
                    CK(cudaStreamWaitEvent(cs, ctx->ev_snap[k], 0));
                    CK(cudaMemcpyAsync(&ctx->h_snap_n[k], ctx->snap.p + k, sizeof(uint32_t), cudaMemcpyDeviceToHost, cs));
                    CK(cudaStreamSynchronize(cs));
                    uint32_t nk = ctx->h_snap_n[k];
                    if (nk > cap_contacts) break;  // the contact array overflowed: this attempt is repeated with a larger one
                    uint32_t wc = nk < ef.cap_contacts ? nk : ef.cap_contacts;
                    if (ef.contacts && wc > ef.contacts_done)
                        CK(cudaMemcpyAsync(ef.contacts + ef.contacts_done, ctx->contacts.p + ef.contacts_done,
                                           sizeof(ncb_contact) * (size_t)(wc - ef.contacts_done), cudaMemcpyDeviceToHost, cs));
                    if (wc > ef.contacts_done) ef.contacts_done = wc;
                }
            }
        }
        r = read_counters(ctx);
        if (r) return r;
        const DevCounters& c = ctx->last_counters;
        bool over = false;
        if (c.n_pairs > cap_pairs) {
            cap_pairs = (size_t)c.n_pairs + c.n_pairs / 8 + 1024;
            over = true;
        }
        if (c.n_contacts > cap_contacts) {
            cap_contacts = (size_t)c.n_contacts + c.n_contacts / 8 + 1024;
            over = true;
        }
        ctx->cap_pairs_hint = (uint32_t)cap_pairs;
        ctx->cap_contacts_hint = (uint32_t)cap_contacts;
        if (!over) {
            ctx->last_n_pairs = c.n_pairs;
            ctx->last_n_contacts = c.n_contacts;
            ctx->have_kinematics = ctx->want_kinematics;
            ctx->early.valid = ctx->early.active;
            ctx->early.active = false;
            return NCB_OK;
        }
    }
    ctx->err = "pair/contact buffers still too small after 3 attempts";
    return NCB_ERR_STATE;
}

int ncb_world_update_stage(ncb_ctx* ctx, int stage, float margin, uint32_t begin, uint32_t end, ncb_update_counts* counts) {
    if (!ctx) return NCB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    uint32_t n = ctx->n;
    if (stage == 0) {
        if (end > n) end = n;
        timer_begin(ctx);
        ctx->timer_external = true;
        CK(launch_aabbs(ctx, dev_objects(ctx), margin, 2, begin, end));
        timer_mark(ctx, "aabb", 1);
        return NCB_OK;
    }
    if (n == 0) {
        memset(&ctx->last_counters, 0, sizeof ctx->last_counters);
        ctx->last_n_pairs = ctx->last_n_contacts = 0;
        fill_counts(ctx, counts);
        return NCB_OK;
    }
    int r = update_after_aabbs(ctx, begin, end);
    ctx->timer_external = false;
    if (r) return r;
    fill_counts(ctx, counts);
    return NCB_OK;
}

// Stage 1 of a multi-GPU update with SPATIAL ownership (see launch_shard_select): every rank holds all fat AABBs (stage 0 +
// all-gather), selects the objects it owns plus the ghosts around them, builds its LBVH over those only and reports the
// pairs it is responsible for.  The union over the ranks is the full pair set, every pair exactly once.
int ncb_world_update_sharded(ncb_ctx* ctx, float margin, int rank, int world, ncb_update_counts* counts) {
    (void)margin;
    if (!ctx || rank < 0 || world < 1 || rank >= world || world > SHARD_MAX_RANKS) return NCB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    uint32_t n = ctx->n;
    if (n == 0 || world == 1) return ncb_world_update_stage(ctx, 1, margin, 0, 0xffffffffu, counts);
    cudaStream_t s = ctx->stream;
    CK(ctx->shard.reserve(1));
    CK(ctx->shard_bins.reserve(n));
    size_t cap = ctx->shard_sel.cap ? ctx->shard_sel.cap : (size_t)n / world + n / 8 + 4096;
    int r = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        CK(ctx->shard_sel.reserve(cap));
        CK(ctx->shard_lo.reserve(cap));
        CK(ctx->shard_hi.reserve(cap));
        cap = std::min(ctx->shard_sel.cap, std::min(ctx->shard_lo.cap, ctx->shard_hi.cap));
        r = reset_counters(ctx);
        if (r) return r;
        if (!ctx->timer_external) timer_begin(ctx);
        CK(launch_shard_select(ctx, n, rank, world, ctx->shard.p, ctx->shard_bins.p, (uint32_t)cap, ctx->shard_sel.p, ctx->shard_lo.p, ctx->shard_hi.p));
        uint32_t mo[2] = {0, 0};
        CK(cudaMemcpyAsync(mo, &ctx->shard.p->m, 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        ctx->shard_m = mo[0], ctx->shard_owned = mo[1];
        if (mo[0] <= cap) break;
        cap = (size_t)mo[0] + mo[0] / 8 + 4096;
    }
    timer_mark(ctx, "shard_select", 5);
    // the selected boxes stand in for the object boxes during the local build / search
    std::swap(ctx->aabb_lo, ctx->shard_lo);
    std::swap(ctx->aabb_hi, ctx->shard_hi);
    ctx->timer_external = true;
    r = update_after_aabbs(ctx, 0, 0xffffffffu, ctx->shard_m, ctx->shard_sel.p, rank);
    ctx->timer_external = false;
    std::swap(ctx->aabb_lo, ctx->shard_lo);
    std::swap(ctx->aabb_hi, ctx->shard_hi);
    if (r) return r;
    fill_counts(ctx, counts);
    return NCB_OK;
}

// Multi-GPU update with ROUTED spatial ownership (see the routed-sharding block of broad.cu): nothing is all-gathered; every rank
// sends each object of its own block [begin, end) to the rank that owns its Morton bin, and as a ghost to the ranks whose region
// its box meets.  The caller performs one collective on the buffers of ncb_route_buffer between the stages:
//   stage 0 (AABBs + centre bounds of the own block)   -> all-reduce MAX  of buffer 0 (6 floats)
//   stage 1 (bins + histogram)                          -> all-reduce SUM  of buffer 1 (SHARD_BINS ints)
//   stage 2 (owner buckets + regions)                   -> all-to-all      buffer 2 -> 3, all-reduce MAX of buffer 4 (8 sub-boxes x 6 floats per rank)
//   stage 3 (ghost buckets)                             -> all-to-all      buffer 5 -> 6
//   stage 4 (unpack, local LBVH, pair search, narrow phase; fills counts).  Returns NCB_ROUTE_REPEAT when a bucket was too
//            small somewhere: the capacities have been raised (identically on every rank), repeat from stage 2.
// with_poses != 0: the records also carry the poses (the end-to-end arm: a rank uploads the poses of its own block only).
static int routed_stage(ncb_ctx* ctx, int stage, float margin, int rank, int world, uint32_t begin, uint32_t end, int with_poses,
                        ncb_update_counts* counts) {
    uint32_t n = ctx->n;
    ncb::RouteBufs& R = ctx->route;
    cudaStream_t s = ctx->stream;
    uint32_t n_own = end - begin;
    const bool p2p = R.p2p;
    if (p2p) {
        REQUIRE(R.p2p_rank == rank && R.p2p_world == world, NCB_ERR_ARG, "ncb_world_update_routed: rank / world differ from ncb_route_p2p_alloc");
        REQUIRE(n_own + 1 <= R.p2p_cap, NCB_ERR_STATE, "ncb_world_update_routed: the block outgrew the peer buffers (ncb_route_p2p_alloc again)");
    }
    if (stage == 0) {
        R.recw = with_poses ? 4 : 2;
        CK(R.bounds.reserve(8));
        CK(R.hist.reserve(SHARD_BINS));
        CK(R.split.reserve(SHARD_MAX_RANKS + 1));
        CK(R.bins.reserve(n_own ? n_own : 1));
        CK(R.region_i.reserve(SHARD_MAX_RANKS * SHARD_SUBS * 6));
        CK(R.region_f.reserve(SHARD_MAX_RANKS * SHARD_SUBS * 6));
        CK(R.counts.reserve(2 * SHARD_MAX_RANKS));
        CK(ctx->shard.reserve(1));
        CK(ctx->counters.reserve(1));
        int r = reset_counters(ctx);
        if (r) return r;
        timer_begin(ctx);
        ctx->timer_external = true;
        CK(launch_aabbs(ctx, dev_objects(ctx), margin, 2, begin, end));
        timer_mark(ctx, "aabb", 1);
        CK(launch_route_stage(ctx, 0, rank, world, begin, end, R));
        if (p2p) {
            R.epoch++;
            CK(launch_p2p_push(ctx, R, R.bounds.p, 6, 0, 1));
        }
        return NCB_OK;
    }
    if (stage == 1) {
        if (p2p) CK(launch_p2p_wait_reduce(ctx, R, 1, 0, 6, 0, R.bounds.p));
        CK(launch_route_stage(ctx, 1, rank, world, begin, end, R));
        if (p2p) CK(launch_p2p_push(ctx, R, R.hist.p, SHARD_BINS, 1, 2));
        return NCB_OK;
    }
    if (stage == 2) {
        if (p2p) {
            CK(launch_p2p_wait_reduce(ctx, R, 2, 1, SHARD_BINS, 1, R.hist.p));
        } else {
            // bucket capacities (records incl. the header slot), the same on every rank: derived from the largest block and from
            // requirements that every rank sees identically (stage 4)
            uint32_t blk = (n + world - 1) / world;
            static const char* slack_env = getenv("NCB_ROUTE_SLACK");  // tests: a tiny slack forces the grow-and-repeat path
            uint32_t slack = slack_env ? (uint32_t)atoi(slack_env) : 2048;
            if (!R.cap_o) R.cap_o = blk / world + blk / (4 * world) + slack;
            if (!R.cap_g) R.cap_g = blk / (4 * world) + slack;
            if (R.cap_o < 2) R.cap_o = 2;
            if (R.cap_g < 2) R.cap_g = 2;
            size_t w = (size_t)R.recw;
            CK(R.send_o.reserve((size_t)world * R.cap_o * w));
            CK(R.recv_o.reserve((size_t)world * R.cap_o * w));
            CK(R.send_g.reserve((size_t)world * R.cap_g * w));
            CK(R.recv_g.reserve((size_t)world * R.cap_g * w));
        }
        CK(launch_route_stage(ctx, 2, rank, world, begin, end, R));
        if (p2p) CK(launch_p2p_push(ctx, R, R.region_f.p, SHARD_SUBS * 6 * (uint32_t)world, 2, 3));
        return NCB_OK;
    }
    if (stage == 3) {
        if (p2p) CK(launch_p2p_wait_reduce(ctx, R, 3, 2, SHARD_SUBS * 6 * (uint32_t)world, 0, R.region_f.p));
        CK(launch_route_stage(ctx, 3, rank, world, begin, end, R));
        if (p2p) CK(launch_p2p_push(ctx, R, nullptr, 0, 2, 4));
        timer_mark(ctx, "route", 5);
        return NCB_OK;
    }
    // stage 4
    if (p2p) CK(launch_p2p_wait_reduce(ctx, R, 4, 0, 0, -1, nullptr));
    uint32_t cap_o = p2p ? R.p2p_cap : R.cap_o, cap_g = p2p ? R.p2p_cap : R.cap_g;
    size_t cap_local = p2p ? (size_t)n : (size_t)world * ((size_t)cap_o + cap_g);  // a rank never holds an object twice
    CK(ctx->shard_sel.reserve(cap_local));
    CK(ctx->shard_lo.reserve(cap_local));
    CK(ctx->shard_hi.reserve(cap_local));
    CK(launch_route_unpack(ctx, world, R, (uint32_t)cap_local, ctx->shard.p, ctx->shard_sel.p, ctx->shard_lo.p, ctx->shard_hi.p));
    uint32_t tail[SHARD_MAX_RANKS + 1 + 6 + 2];
    uint32_t p2p_err = 0;
    CK(cudaMemcpyAsync(tail, &ctx->shard.p->split[0], sizeof tail, cudaMemcpyDeviceToHost, s));
    if (p2p) CK(cudaMemcpyAsync(&p2p_err, R.p2p_meta.p + P2P_ERR_OFF, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (p2p_err) {
        ctx->timer_external = false;
        ctx->err = "routed update: a peer did not reach the exchange within the time limit";
        return NCB_ERR_STATE;
    }
    uint32_t need_o = tail[0], need_g = tail[1];
    ctx->shard_m = tail[SHARD_MAX_RANKS + 1 + 6], ctx->shard_owned = tail[SHARD_MAX_RANKS + 1 + 6 + 1];
    if (need_o > cap_o || need_g > cap_g) {
        if (p2p) {
            ctx->err = "routed update: a peer bucket overflowed";
            return NCB_ERR_STATE;
        }
        if (need_o > R.cap_o) R.cap_o = need_o + need_o / 8 + 1024;
        if (need_g > R.cap_g) R.cap_g = need_g + need_g / 8 + 1024;
        return NCB_ROUTE_REPEAT;
    }
    timer_mark(ctx, "route_unpack", 7);
    if (ctx->shard_m == 0) {
        ctx->timer_external = false;
        memset(&ctx->last_counters, 0, sizeof ctx->last_counters);
        ctx->last_n_pairs = ctx->last_n_contacts = 0;
        ctx->early.active = ctx->early.valid = false;
        fill_counts(ctx, counts);
        return NCB_OK;
    }
    std::swap(ctx->aabb_lo, ctx->shard_lo);
    std::swap(ctx->aabb_hi, ctx->shard_hi);
    int r = update_after_aabbs(ctx, 0, 0xffffffffu, ctx->shard_m, ctx->shard_sel.p, rank);
    ctx->timer_external = false;
    std::swap(ctx->aabb_lo, ctx->shard_lo);
    std::swap(ctx->aabb_hi, ctx->shard_hi);
    if (r) return r;
    fill_counts(ctx, counts);
    return NCB_OK;
}

int ncb_world_update_routed(ncb_ctx* ctx, int stage, float margin, int rank, int world, uint32_t begin, uint32_t end, int with_poses,
                            ncb_update_counts* counts) {
    if (!ctx || rank < 0 || world < 1 || rank >= world || world > SHARD_MAX_RANKS || stage < -1 || stage > 4) return NCB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (end > ctx->n) end = ctx->n;
    if (begin > end) begin = end;
    if (stage >= 0) return routed_stage(ctx, stage, margin, rank, world, begin, end, with_poses, counts);
    // stage -1: the whole step in one call; only with the peer-memory exchange (nothing for the caller to do between the stages)
    REQUIRE(ctx->route.p2p, NCB_ERR_STATE, "ncb_world_update_routed(stage -1) needs ncb_route_p2p_connect");
    for (int st = 0; st <= 4; ++st) {
        int r = routed_stage(ctx, st, margin, rank, world, begin, end, with_poses, counts);
        if (r) return r;
    }
    return NCB_OK;
}

// ---- peer-memory exchange set-up -------------------------------------------------------------------------------------------
// Allocates this rank's receive buffers (owner buckets, ghost buckets: world x (largest block + 1) records of 64 B, so they can
// never overflow) and its meta / flag words, and exports them: `handles` receives 3 cudaIpcMemHandle_t (3 x 64 B) for peers in other
// processes, `ptrs` the 3 raw device pointers for peers inside this process.
int ncb_route_p2p_alloc(ncb_ctx* ctx, int rank, int world, uint32_t n_total, void* handles, uint64_t* ptrs) {
    if (!ctx || rank < 0 || world < 1 || rank >= world || world > SHARD_MAX_RANKS) return NCB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    ncb::RouteBufs& R = ctx->route;
    uint32_t blk = (n_total + world - 1) / world;
    R.p2p = false;
    R.p2p_rank = rank, R.p2p_world = world, R.p2p_cap = blk + 1, R.epoch = 0;
    size_t recs = (size_t)world * R.p2p_cap * 4;
    CK(R.p2p_recv_o.reserve(recs));
    CK(R.p2p_recv_g.reserve(recs));
    CK(R.p2p_meta.reserve(P2P_META_WORDS));
    CK(cudaMemsetAsync(R.p2p_meta.p, 0, P2P_META_WORDS * sizeof(uint32_t), ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    void* base[3] = {R.p2p_recv_o.p, R.p2p_recv_g.p, R.p2p_meta.p};
    for (int k = 0; k < 3; ++k) {
        if (ptrs) ptrs[k] = (uint64_t)(uintptr_t)base[k];
        if (handles) CK(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handles + k, base[k]));
    }
    return NCB_OK;
}

static void p2p_close(ncb_ctx* ctx) {
    ncb::RouteBufs& R = ctx->route;
    for (int q = 0; q < SHARD_MAX_RANKS; ++q) {
        void* mapped[3] = {R.peer_recv_o[q], R.peer_recv_g[q], R.peer_meta[q]};
        for (int k = 0; k < 3; ++k)
            if (R.peer_opened[q][k] && mapped[k]) cudaIpcCloseMemHandle(mapped[k]);
        R.peer_recv_o[q] = R.peer_recv_g[q] = nullptr;
        R.peer_meta[q] = nullptr;
        R.peer_opened[q][0] = R.peer_opened[q][1] = R.peer_opened[q][2] = false;
    }
    R.p2p = false;
}

// Connects to the peers: handles_all = world x 3 cudaIpcMemHandle_t in rank order (peers in other processes), or NULL with ptrs_all =
// world x 3 raw device pointers (all ranks live in this process: the replay tests).  The own rank always uses its local pointers.
int ncb_route_p2p_connect(ncb_ctx* ctx, const void* handles_all, const uint64_t* ptrs_all) {
    if (!ctx || (!handles_all && !ptrs_all)) return NCB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    ncb::RouteBufs& R = ctx->route;
    REQUIRE(R.p2p_world > 0 && R.p2p_meta.p, NCB_ERR_STATE, "ncb_route_p2p_connect: call ncb_route_p2p_alloc first");
    p2p_close(ctx);
    for (int q = 0; q < R.p2p_world; ++q) {
        void* got[3] = {nullptr, nullptr, nullptr};
        if (q == R.p2p_rank) {
            got[0] = R.p2p_recv_o.p, got[1] = R.p2p_recv_g.p, got[2] = R.p2p_meta.p;
        } else if (handles_all) {
            for (int k = 0; k < 3; ++k) {
                cudaError_t e = cudaIpcOpenMemHandle(&got[k], ((const cudaIpcMemHandle_t*)handles_all)[q * 3 + k], cudaIpcMemLazyEnablePeerAccess);
                if (e != cudaSuccess) {
                    ctx->err = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e);
                    p2p_close(ctx);
                    return NCB_ERR_CUDA;
                }
                R.peer_opened[q][k] = true;
                if (k == 0) R.peer_recv_o[q] = (float4*)got[0];  // recorded at once so that a later failure closes it
                if (k == 1) R.peer_recv_g[q] = (float4*)got[1];
                if (k == 2) R.peer_meta[q] = (uint32_t*)got[2];
            }
        } else {
            for (int k = 0; k < 3; ++k) got[k] = (void*)(uintptr_t)ptrs_all[q * 3 + k];
        }
        R.peer_recv_o[q] = (float4*)got[0], R.peer_recv_g[q] = (float4*)got[1], R.peer_meta[q] = (uint32_t*)got[2];
    }
    R.p2p = true;
    return NCB_OK;
}

int ncb_route_p2p_close(ncb_ctx* ctx) {
    if (!ctx) return NCB_ERR_ARG;
    cudaSetDevice(ctx->device);
    p2p_close(ctx);
    return NCB_OK;
}

// Buffers of the routed update for the caller's collectives: which = 0 bounds (6 f32), 1 histogram (SHARD_BINS i32), 2 / 3 owner
// buckets send / recv, 4 regions (8 sub-boxes x 6 f32 per rank), 5 / 6 ghost buckets send / recv.
// *bytes = the extent a collective covers (for 2 / 3 / 5 / 6: world equal parts).  Valid after the stage that precedes the collective.
void* ncb_route_buffer(ncb_ctx* ctx, int which, int world, uint64_t* bytes) {
    if (!ctx) return nullptr;
    ncb::RouteBufs& R = ctx->route;
    uint64_t b = 0;
    void* p = nullptr;
    switch (which) {
        case 0: p = R.bounds.p, b = 6 * sizeof(float); break;
        case 1: p = R.hist.p, b = SHARD_BINS * sizeof(int); break;
        case 2: p = R.send_o.p, b = (uint64_t)world * R.cap_o * R.recw * sizeof(float4); break;
        case 3: p = R.recv_o.p, b = (uint64_t)world * R.cap_o * R.recw * sizeof(float4); break;
        case 4: p = R.region_f.p, b = (uint64_t)world * SHARD_SUBS * 6 * sizeof(float); break;
        case 5: p = R.send_g.p, b = (uint64_t)world * R.cap_g * R.recw * sizeof(float4); break;
        case 6: p = R.recv_g.p, b = (uint64_t)world * R.cap_g * R.recw * sizeof(float4); break;
        default: break;
    }
    if (bytes) *bytes = b;
    return p;
}

int ncb_world_update_device(ncb_ctx* ctx, float margin, uint32_t q_begin, uint32_t q_end, ncb_update_counts* counts) {
    if (!ctx) return NCB_ERR_ARG;
    int r = ncb_world_update_stage(ctx, 0, margin, 0, ctx->n, nullptr);
    if (r) return r;
    return ncb_world_update_stage(ctx, 1, margin, q_begin, q_end, counts);
}

// Arms the overlapped result fetch for the NEXT device update (ncb_world_update_device / _stage(1) / _sharded): while its
// narrow phase runs, the sorted pairs (+ algorithm) and the contacts that are already final are copied into these host
// buffers; ncb_world_fetch with the same buffers then copies only the rest.  ncb_world_update does this by itself.
int ncb_world_fetch_early(ncb_ctx* ctx, uint32_t* pairs, uint32_t cap_pairs, uint8_t* pair_algo, ncb_contact* contacts, uint32_t cap_contacts) {
    if (!ctx) return NCB_ERR_ARG;
    static const bool early_ok = getenv("NCB_NO_EARLY_FETCH") == nullptr;
    ctx->early.active = early_ok && ctx->copy_stream != nullptr;
    ctx->early.pairs = pairs, ctx->early.algo = pair_algo, ctx->early.contacts = contacts;
    ctx->early.cap_pairs = cap_pairs, ctx->early.cap_contacts = cap_contacts;
    ctx->early.pairs_done = ctx->early.contacts_done = 0;
    ctx->early.valid = false;
    return NCB_OK;
}

int ncb_world_fetch(ncb_ctx* ctx, uint32_t* pairs, uint32_t cap_pairs, uint8_t* pair_algo, uint32_t* manifold_start,
                    uint8_t* manifold_count, ncb_contact* contacts, uint32_t cap_contacts) {
    if (!ctx) return NCB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    uint32_t np = ctx->last_n_pairs, nc = ctx->last_n_contacts;
    uint32_t wp = np < cap_pairs ? np : cap_pairs, wc = nc < cap_contacts ? nc : cap_contacts;
    // rows an armed early fetch already put on their way into exactly these buffers
    uint32_t pd = 0, cd = 0;
    if (ctx->early.valid) {
        if (pairs == ctx->early.pairs && pair_algo == ctx->early.algo) pd = ctx->early.pairs_done;
        if (contacts == ctx->early.contacts) cd = ctx->early.contacts_done;
        ctx->early.valid = false;
    }
    ctx->early.active = false;
    if (pd > wp) pd = wp;
    if (cd > wc) cd = wc;
    if (pairs && wp > pd) CK(cudaMemcpyAsync(pairs + 2 * (size_t)pd, ctx->pairs.p + pd, 8 * (size_t)(wp - pd), cudaMemcpyDeviceToHost, s));
    if (pair_algo && wp > pd) CK(cudaMemcpyAsync(pair_algo + pd, ctx->pair_algo.p + pd, wp - pd, cudaMemcpyDeviceToHost, s));
    if (manifold_start && wp) CK(cudaMemcpyAsync(manifold_start, ctx->manifold_start.p, 4 * (size_t)wp, cudaMemcpyDeviceToHost, s));
    if (manifold_count && wp) CK(cudaMemcpyAsync(manifold_count, ctx->manifold_count.p, wp, cudaMemcpyDeviceToHost, s));
    if (contacts && wc > cd) CK(cudaMemcpyAsync(contacts + cd, ctx->contacts.p + cd, sizeof(ncb_contact) * (size_t)(wc - cd), cudaMemcpyDeviceToHost, s));
    if (ctx->copy_stream) CK(cudaStreamSynchronize(ctx->copy_stream));
    CK(cudaStreamSynchronize(s));
    return ((pairs && np > cap_pairs) || (contacts && nc > cap_contacts)) ? 1 : NCB_OK;
}

int ncb_world_fetch_proximity(ncb_ctx* ctx, uint8_t* prox, uint32_t cap_pairs) {
    if (!ctx || (cap_pairs && !prox)) return NCB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    uint32_t np = ctx->last_n_pairs;
    uint32_t wp = np < cap_pairs ? np : cap_pairs;
    if (wp == 0) return np > cap_pairs ? 1 : NCB_OK;
    if (!ctx->has_prox || !ctx->prox.p) {
        memset(prox, 0xff, wp);
    } else {
        CK(cudaMemcpyAsync(prox, ctx->prox.p, wp, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return np > cap_pairs ? 1 : NCB_OK;
}

int ncb_world_update(ncb_ctx* ctx, const ncb_objects* objs, float margin, uint32_t* pairs, uint32_t cap_pairs, uint8_t* pair_algo,
                     uint32_t* manifold_start, uint8_t* manifold_count, ncb_contact* contacts, uint32_t cap_contacts,
                     ncb_update_counts* counts) {
    if (!ctx || !objs) return NCB_ERR_ARG;
    // the query types installed with ncb_set_query_types survive this call when the object count is unchanged (the call re-uploads
    // the world it was given before); ncb_set_objects alone resets them
    const bool keep_kinds = ctx->has_prox && objs->n == ctx->n;
    int r = ncb_set_objects(ctx, objs);
    if (r) return r;
    if (keep_kinds) ctx->has_prox = true;
    // results are copied back while the narrow phase is still running (see update_after_aabbs); NCB_NO_EARLY_FETCH=1 disables it
    r = ncb_world_fetch_early(ctx, pairs, cap_pairs, pair_algo, contacts, cap_contacts);
    if (r) return r;
    r = ncb_world_update_device(ctx, margin, 0, 0xffffffffu, counts);
    if (r) {
        ctx->early.active = ctx->early.valid = false;
        if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
        return r;
    }
    return ncb_world_fetch(ctx, pairs, cap_pairs, pair_algo, manifold_start, manifold_count, contacts, cap_contacts);
}

int ncb_world_update_poses(ncb_ctx* ctx, uint32_t n, const float* pos, const float* rot, float margin, uint32_t* pairs, uint32_t cap_pairs,
                           uint8_t* pair_algo, uint32_t* manifold_start, uint8_t* manifold_count, ncb_contact* contacts,
                           uint32_t cap_contacts, ncb_update_counts* counts) {
    if (!ctx || (n && (!pos || !rot))) return NCB_ERR_ARG;
    int r = ncb_set_positions(ctx, n, pos, rot);
    if (r) return r;
    r = ncb_world_fetch_early(ctx, pairs, cap_pairs, pair_algo, contacts, cap_contacts);
    if (r) return r;
    r = ncb_world_update_device(ctx, margin, 0, 0xffffffffu, counts);
    if (r) {
        ctx->early.active = ctx->early.valid = false;
        if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
        return r;
    }
    return ncb_world_fetch(ctx, pairs, cap_pairs, pair_algo, manifold_start, manifold_count, contacts, cap_contacts);
}

void* ncb_device_ptr(ncb_ctx* ctx, int which) {
    if (!ctx) return nullptr;
    switch (which) {
        case 0: return ctx->aabb_lo.p;
        case 1: return ctx->aabb_hi.p;
        case 2: return ctx->pairs.p;
        case 3: return ctx->contacts.p;
        case 4: return ctx->pos.p;
        case 5: return ctx->rot.p;
        default: return nullptr;
    }
}

}  // extern "C"
