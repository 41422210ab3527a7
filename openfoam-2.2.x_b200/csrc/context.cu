// C-ABI entry points: library state, context, device memory, matrix objects.
#include <algorithm>
#include <cstring>
#include <thread>

#include "ldu_internal.h"
#include "sweeps.h"

namespace ldu {

long long g_launches = 0;
static thread_local std::string g_error;

void set_error(const std::string& msg) { g_error = msg; }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line)
{
    char buf[1024];
    snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e),
             file, line, what);
    g_error = buf;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInitializationError)
        return LDU_ENODEVICE;
    return LDU_ECUDA;
}

template <class T>
static int upload(ldu_context* ctx, T** d, const T* h, size_t n, size_t pad = 0)
{
    // pad: zeroed elements behind the array (the TMA-staged row kernel copies 16-byte aligned
    // ranges that may reach a few elements past the end)
    *d = nullptr;
    LDU_CUDA(cudaMalloc((void**)d, std::max<size_t>(n + pad, 1) * sizeof(T)));
    if (pad) LDU_CUDA(cudaMemsetAsync(*d + n, 0, pad * sizeof(T), ctx->stream));
    if (n) LDU_CUDA(cudaMemcpyAsync(*d, h, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return LDU_OK;
}

// ---------------------------------------------------------------------------
// large copies from / to pageable host memory: several host threads stage chunks through pinned buffers
// ---------------------------------------------------------------------------
namespace {
constexpr int kStageThreads = 12;    // upper bound; LDU_STAGE_THREADS picks fewer (default 8)
constexpr size_t kStageChunk = 4u << 20;            // size of a pinned staging buffer
constexpr size_t kStageChunkDefault = 4u << 20;     // bytes staged per copy
constexpr size_t kStageMin = 1u << 20;     // below this a plain copy is as fast

struct StageWorker {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    unsigned char* buf[2] = {nullptr, nullptr};
};
struct StagePool {
    StageWorker w[kStageThreads];
};

bool is_pageable(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

int stage_pool(ldu_context* ctx, StagePool** out)
{
    if (!ctx->stage) {
        StagePool* sp = new StagePool();
        ctx->stage = sp;
        for (int t = 0; t < kStageThreads; t++) {
            LDU_CUDA(cudaStreamCreateWithFlags(&sp->w[t].stream, cudaStreamNonBlocking));
            for (int b = 0; b < 2; b++) {
                LDU_CUDA(cudaEventCreateWithFlags(&sp->w[t].ev[b], cudaEventDisableTiming));
                LDU_CUDA(cudaMallocHost((void**)&sp->w[t].buf[b], kStageChunk));
            }
        }
    }
    *out = reinterpret_cast<StagePool*>(ctx->stage);
    return LDU_OK;
}

// chunk k of the copy belongs to worker k % kStageThreads, buffer (k / kStageThreads) % 2
int staged_copy(ldu_context* ctx, unsigned char* dst, const unsigned char* src, size_t bytes, bool toDevice)
{
    StagePool* sp = nullptr;
    LDU_TRY(stage_pool(ctx, &sp));
    LDU_CUDA(cudaStreamSynchronize(ctx->stream));     // ordered after what is queued on the context's stream
    // LDU_STAGE_CHUNK_KB: bytes staged per copy (at most the size of the pinned buffers)
    static const size_t chunk = [] {
        const char* e = getenv("LDU_STAGE_CHUNK_KB");
        const size_t v = e ? (size_t)atol(e) << 10 : kStageChunkDefault;
        return std::max<size_t>(64u << 10, std::min(v, kStageChunk));
    }();
    const size_t nChunks = (bytes + chunk - 1) / chunk;
    static const int wanted = [] {
        const char* e = getenv("LDU_STAGE_THREADS");
        const int hw = (int)std::thread::hardware_concurrency();
        int n = e ? atoi(e) : 8;
        if (hw > 0) n = std::min(n, std::max(hw - 1, 1));
        return std::max(1, std::min(n, kStageThreads));
    }();
    const int nThreads = (int)std::min<size_t>((size_t)wanted, nChunks);
    std::vector<cudaError_t> err(nThreads, cudaSuccess);
    auto work = [&](int t) {
        cudaSetDevice(ctx->device);
        StageWorker& w = sp->w[t];
        cudaError_t e = cudaSuccess;
        size_t pending[2] = {0, 0}, pendingOff[2] = {0, 0};      // D2H: bytes waiting in buffer b
        int use = 0;
        for (size_t k = t; k < nChunks && e == cudaSuccess; k += nThreads, use ^= 1) {
            const size_t off = k * chunk, n = std::min(chunk, bytes - off);
            if (toDevice) {
                e = cudaEventSynchronize(w.ev[use]);             // the copy that last used this buffer is done
                if (e != cudaSuccess) break;
                memcpy(w.buf[use], src + off, n);
                e = cudaMemcpyAsync(dst + off, w.buf[use], n, cudaMemcpyHostToDevice, w.stream);
                if (e == cudaSuccess) e = cudaEventRecord(w.ev[use], w.stream);
            } else {
                if (pending[use]) {                              // drain the buffer before reusing it
                    e = cudaEventSynchronize(w.ev[use]);
                    if (e != cudaSuccess) break;
                    memcpy(dst + pendingOff[use], w.buf[use], pending[use]);
                }
                e = cudaMemcpyAsync(w.buf[use], src + off, n, cudaMemcpyDeviceToHost, w.stream);
                if (e == cudaSuccess) e = cudaEventRecord(w.ev[use], w.stream);
                pending[use] = n;
                pendingOff[use] = off;
            }
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(w.stream);
        if (!toDevice && e == cudaSuccess)
            for (int b = 0; b < 2; b++) {
                const int q = use ^ b;                           // older buffer first (order is irrelevant, both are complete)
                if (pending[q]) memcpy(dst + pendingOff[q], w.buf[q], pending[q]);
            }
        err[t] = e;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nThreads; t++) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    for (cudaError_t e : err)
        if (e != cudaSuccess) return cuda_fail(e, "staged copy", __FILE__, __LINE__);
    return LDU_OK;
}
}  // namespace

int copy_h2d(ldu_context* ctx, void* dst, const void* src, size_t bytes)
{
    if (!bytes) return LDU_OK;
    if (bytes >= kStageMin && is_pageable(src))
        return staged_copy(ctx, (unsigned char*)dst, (const unsigned char*)src, bytes, true);
    LDU_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    LDU_CUDA(cudaStreamSynchronize(ctx->stream));
    return LDU_OK;
}

int copy_d2h(ldu_context* ctx, void* dst, const void* src, size_t bytes)
{
    if (!bytes) return LDU_OK;
    if (bytes >= kStageMin && is_pageable(dst))
        return staged_copy(ctx, (unsigned char*)dst, (const unsigned char*)src, bytes, false);
    LDU_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    LDU_CUDA(cudaStreamSynchronize(ctx->stream));
    return LDU_OK;
}

void stage_free(ldu_context* ctx)
{
    StagePool* sp = reinterpret_cast<StagePool*>(ctx->stage);
    if (!sp) return;
    for (int t = 0; t < kStageThreads; t++) {
        if (sp->w[t].stream) cudaStreamDestroy(sp->w[t].stream);
        for (int b = 0; b < 2; b++) {
            if (sp->w[t].ev[b]) cudaEventDestroy(sp->w[t].ev[b]);
            if (sp->w[t].buf[b]) cudaFreeHost(sp->w[t].buf[b]);
        }
    }
    delete sp;
    ctx->stage = nullptr;
}

double* work_vec(ldu_matrix* m, int idx)
{
    if ((int)m->work.size() <= idx) m->work.resize(idx + 1, nullptr);
    if (!m->work[idx]) {
        if (cudaMalloc((void**)&m->work[idx], std::max(m->nCells, 1) * sizeof(double)) != cudaSuccess)
            return nullptr;
    }
    return m->work[idx];
}

int ensure_scalars(ldu_matrix* m)
{
    if (!m->d_scalars) {
        LDU_CUDA(cudaMalloc((void**)&m->d_scalars, sizeof(SolverScalars)));
        LDU_CUDA(cudaMemsetAsync(m->d_scalars, 0, sizeof(SolverScalars), m->ctx->stream));
    }
    if (!m->d_hist) LDU_CUDA(cudaMalloc((void**)&m->d_hist, kMaxHist * sizeof(double)));
    return LDU_OK;
}

// greedy first-fit colouring of the cell graph in cell order (ldu_colour_order, multiColourGaussSeidel)
int greedy_colouring(int nCells, int nFaces, const int* lowerAddr, const int* upperAddr, std::vector<int>& colour,
                     int* nColours)
{
    // CSR adjacency (both directions)
    std::vector<int> start(nCells + 1, 0);
    for (int f = 0; f < nFaces; f++) {
        const int l = lowerAddr[f], u = upperAddr[f];
        if (l < 0 || u < 0 || l >= nCells || u >= nCells || l == u) {
            set_error("colouring: addressing out of range");
            return LDU_EINVAL;
        }
        start[l + 1]++;
        start[u + 1]++;
    }
    for (int c = 0; c < nCells; c++) start[c + 1] += start[c];
    std::vector<int> adj(start[nCells]), fill(start.begin(), start.end() - 1);
    for (int f = 0; f < nFaces; f++) {
        adj[fill[lowerAddr[f]]++] = upperAddr[f];
        adj[fill[upperAddr[f]]++] = lowerAddr[f];
    }
    colour.assign(nCells, -1);
    std::vector<int> mark;
    int nc = 0;
    for (int c = 0; c < nCells; c++) {
        mark.assign(nc + 1, 0);
        for (int k = start[c]; k < start[c + 1]; k++) {
            const int q = colour[adj[k]];
            if (q >= 0) mark[q] = 1;
        }
        int q = 0;
        while (q < nc && mark[q]) q++;
        colour[c] = q;
        if (q == nc) nc++;
    }
    *nColours = nc;
    return LDU_OK;
}

}  // namespace ldu

using namespace ldu;

extern "C" {

const char* ldu_version(void) { return "ldu_b200 0.1 (sm_100a)"; }
const char* ldu_last_error(void) { return g_error.c_str(); }
long long ldu_launch_count(void) { return g_launches; }

void ldu_controls_default(ldu_controls* c)
{
    memset(c, 0, sizeof(*c));
    c->solver = LDU_SOLVER_PCG;
    c->preconditioner = LDU_PRECOND_NONE;
    c->smoother = LDU_SMOOTHER_GS;
    c->maxIter = 1000;
    c->tolerance = 1e-6;
    c->relTol = 0.0;
    c->nSweeps = 1;
    c->nCellsInCoarsestLevel = 10;
    c->mergeLevels = 1;
    c->nPreSweeps = 0;
    c->preSweepsLevelMultiplier = 1;
    c->maxPreSweeps = 4;
    c->nPostSweeps = 2;
    c->postSweepsLevelMultiplier = 1;
    c->maxPostSweeps = 4;
    c->nFinestSweeps = 2;
    c->interpolateCorrection = 0;
    c->scaleCorrection = -1;
    c->nVcycles = 2;
    c->precTolerance = 1e-6;
    c->precRelTol = 0.0;
    c->useFaceWeights = 0;
    c->cacheAgglomeration = 1;
    c->checkInterval = 0;
}

int ldu_device_count(void)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) return 0;
    return count;
}

int ldu_context_create(int device, void* stream, ldu_context** out)
{
    if (!out) return LDU_EINVAL;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("ldu_context_create: no CUDA device available (there is no CPU fallback)");
        return LDU_ENODEVICE;
    }
    if (device < 0 || device >= count) {
        set_error("ldu_context_create: bad device ordinal");
        return LDU_EINVAL;
    }
    LDU_CUDA(cudaSetDevice(device));
    ldu_context* ctx = new ldu_context();
    ctx->device = device;
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        LDU_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->ownStream = true;
    }
    cudaDeviceProp prop;
    LDU_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->smCount = prop.multiProcessorCount;
    ctx->maxBlocks = ctx->smCount * 32;
    LDU_CUDA(cudaMalloc((void**)&ctx->d_partials, (size_t)ctx->maxBlocks * kMaxRed * sizeof(double)));
    LDU_CUDA(cudaMalloc((void**)&ctx->d_ticket, 64));
    LDU_CUDA(cudaMemsetAsync(ctx->d_ticket, 0, 64, ctx->stream));
    LDU_CUDA(cudaMalloc((void**)&ctx->d_red, kMaxRed * sizeof(double)));
    LDU_CUDA(cudaMallocHost((void**)&ctx->h_scalars, sizeof(SolverScalars)));
    *out = ctx;
    return LDU_OK;
}

int ldu_context_destroy(ldu_context* ctx)
{
    if (!ctx) return LDU_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_partials);
    cudaFree(ctx->d_ticket);
    cudaFree(ctx->d_red);
    cudaFreeHost(ctx->h_scalars);
    for (int r = 0; r < ctx->comm.nRanks; r++) {
        if (ctx->comm.connected && r != ctx->comm.rank && ctx->comm.peer[r])
            cudaIpcCloseMemHandle(ctx->comm.peer[r]);
    }
    cudaFree(ctx->comm.d_peer);
    cudaFree(ctx->comm.window);
    stage_free(ctx);
    if (ctx->ownStream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return LDU_OK;
}

int ldu_context_synchronize(ldu_context* ctx)
{
    if (!ctx) return LDU_EINVAL;
    LDU_CUDA(cudaStreamSynchronize(ctx->stream));
    return LDU_OK;
}

void* ldu_context_stream(ldu_context* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int ldu_device_alloc(ldu_context* ctx, long long bytes, void** dptr)
{
    if (!ctx || !dptr || bytes < 0) return LDU_EINVAL;
    LDU_CUDA(cudaSetDevice(ctx->device));
    LDU_CUDA(cudaMalloc(dptr, (size_t)std::max<long long>(bytes, 8)));
    return LDU_OK;
}

int ldu_device_free(ldu_context* ctx, void* dptr)
{
    if (!ctx) return LDU_EINVAL;
    LDU_CUDA(cudaStreamSynchronize(ctx->stream));
    LDU_CUDA(cudaFree(dptr));
    return LDU_OK;
}

int ldu_copy_h2d(ldu_context* ctx, void* dst, const void* src, long long bytes)
{
    if (!ctx) return LDU_EINVAL;
    LDU_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, ctx->stream));
    return LDU_OK;
}

int ldu_copy_d2h(ldu_context* ctx, void* dst, const void* src, long long bytes)
{
    if (!ctx) return LDU_EINVAL;
    LDU_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
    LDU_CUDA(cudaStreamSynchronize(ctx->stream));
    return LDU_OK;
}

int ldu_device_memset(ldu_context* ctx, void* dptr, int byteValue, long long bytes)
{
    if (!ctx) return LDU_EINVAL;
    LDU_CUDA(cudaMemsetAsync(dptr, byteValue, (size_t)bytes, ctx->stream));
    return LDU_OK;
}

int ldu_host_alloc(long long bytes, void** hptr)
{
    if (!hptr) return LDU_EINVAL;
    LDU_CUDA(cudaMallocHost(hptr, (size_t)std::max<long long>(bytes, 8)));
    return LDU_OK;
}

int ldu_host_free(void* hptr)
{
    LDU_CUDA(cudaFreeHost(hptr));
    return LDU_OK;
}

// ---------------------------------------------------------------------------
// matrix
// ---------------------------------------------------------------------------

// Row views of the LDU addressing (lduAddressing.C:31-169).  ownerStart is the
// reference's; losortStart is a proper CSR pointer (the reference's version
// leaves trailing entries at 0, lduAddressing.C:126-169, and never reads them).
// Is this the addressing of a lexicographic nx*ny*nz hex box with faces listed per
// cell in the order +i, +j, +k (blockMesh + fvMeshLduAddressing)?  Decided from the
// owner/neighbour lists alone.
static void detect_box(ldu_matrix* m)
{
    const int n = m->nCells, nf = m->nFaces;
    m->box[0] = m->box[1] = m->box[2] = 0;
    if (n < 1) return;
    // distinct strides u - l, at most three: 1, nx, nx*ny
    long long strides[3] = {0, 0, 0};
    int ns = 0;
    for (int f = 0; f < nf; f++) {
        const long long d = (long long)m->h_u[f] - m->h_l[f];
        bool seen = false;
        for (int q = 0; q < ns; q++) seen = seen || strides[q] == d;
        if (!seen) {
            if (ns == 3) return;
            strides[ns++] = d;
        }
    }
    std::sort(strides, strides + ns);
    long long nx = n, ny = 1, nz = 1;
    if (ns >= 1 && strides[0] != 1) {
        // no i-faces: a single column of cells in i (nx == 1) is not handled here
        return;
    }
    if (ns >= 2) {
        nx = strides[1];
        if (n % nx) return;
        ny = n / nx;
    }
    if (ns == 3) {
        if (strides[2] % nx) return;
        ny = strides[2] / nx;
        if (n % (nx * ny)) return;
        nz = n / (nx * ny);
    }
    if (nx * ny * nz != n) return;
    long long expect = (nx - 1) * ny * nz + nx * (ny - 1) * nz + nx * ny * (nz - 1);
    if (expect != nf) return;
    // every cell owns exactly the faces +i, +j, +k that exist, in that order
    int f = 0;
    for (long long c = 0; c < n; c++) {
        const long long i = c % nx, j = (c / nx) % ny, k = c / (nx * ny);
        if (i < nx - 1) {
            if (f >= nf || m->h_l[f] != c || m->h_u[f] != c + 1) return;
            f++;
        }
        if (j < ny - 1) {
            if (f >= nf || m->h_l[f] != c || m->h_u[f] != c + nx) return;
            f++;
        }
        if (k < nz - 1) {
            if (f >= nf || m->h_l[f] != c || m->h_u[f] != c + nx * ny) return;
            f++;
        }
    }
    if (f != nf) return;
    m->box[0] = (int)nx;
    m->box[1] = (int)ny;
    m->box[2] = (int)nz;
}

static void build_row_views(ldu_matrix* m)
{
    const int n = m->nCells, nf = m->nFaces;
    m->h_ownerStart.assign(n + 1, 0);
    m->h_losortStart.assign(n + 1, 0);
    m->h_losort.assign(nf, 0);
    for (int f = 0; f < nf; f++) {
        m->h_ownerStart[m->h_l[f] + 1]++;
        m->h_losortStart[m->h_u[f] + 1]++;
    }
    for (int c = 0; c < n; c++) {
        m->h_ownerStart[c + 1] += m->h_ownerStart[c];
        m->h_losortStart[c + 1] += m->h_losortStart[c];
    }
    std::vector<int> fill(m->h_losortStart.begin(), m->h_losortStart.end() - 1);
    for (int f = 0; f < nf; f++) m->h_losort[fill[m->h_u[f]]++] = f;  // ascending f per cell
    // are the faces of every owner also sorted by neighbour?  True for meshes in upper-triangular
    // order; GAMG's coarse addressing numbers the faces of a coarse cell in order of discovery
    // (GAMGAgglomerateLduAddressing.C:116-190), where a walk in losort order and a walk in face order
    // visit the faces of one owner differently
    m->nbrSorted = true;
    for (int f = 1; f < nf && m->nbrSorted; f++)
        if (m->h_l[f] == m->h_l[f - 1] && m->h_u[f] < m->h_u[f - 1]) m->nbrSorted = false;
}

int ldu_matrix_create(ldu_context* ctx, int nCells, int nFaces, const int* lowerAddr,
                      const int* upperAddr, int nInterfaces, const int* ifaceSizes,
                      const int* const* faceCells, const int* nbrRank, const int* nbrInterface,
                      ldu_matrix** out)
{
    if (!ctx || !out || nCells < 0 || nFaces < 0 || (nFaces && (!lowerAddr || !upperAddr))) {
        set_error("ldu_matrix_create: bad argument");
        return LDU_EINVAL;
    }
    LDU_CUDA(cudaSetDevice(ctx->device));
    // LDU invariants (lduAddressing.H:36-63): l < u, faces sorted by owner
    for (int f = 0; f < nFaces; f++) {
        const int l = lowerAddr[f], u = upperAddr[f];
        if (l < 0 || u >= nCells || l >= u || (f && lowerAddr[f - 1] > l)) {
            set_error("ldu_matrix_create: addressing is not in upper-triangular order");
            return LDU_EINVAL;
        }
    }
    ldu_matrix* m = new ldu_matrix();
    m->ctx = ctx;
    // any failure below releases what has been built so far (ldu_matrix_destroy copes with a half-built matrix)
    struct Guard {
        ldu_matrix* m;
        ~Guard() { if (m) ldu_matrix_destroy(m); }
    } guard{m};
    m->nCells = nCells;
    m->nFaces = nFaces;
    m->h_l.assign(lowerAddr, lowerAddr + nFaces);
    m->h_u.assign(upperAddr, upperAddr + nFaces);
    build_row_views(m);
    detect_box(m);
    std::vector<int> lowerCol(nFaces);
    for (int k = 0; k < nFaces; k++) lowerCol[k] = m->h_l[m->h_losort[k]];

    LDU_TRY(upload(ctx, &m->d_l, m->h_l.data(), nFaces));
    LDU_TRY(upload(ctx, &m->d_u, m->h_u.data(), nFaces, 8));
    LDU_TRY(upload(ctx, &m->d_ownerStart, m->h_ownerStart.data(), nCells + 1, kRowBlock + 8));
    LDU_TRY(upload(ctx, &m->d_losortStart, m->h_losortStart.data(), nCells + 1, kRowBlock + 8));
    LDU_TRY(upload(ctx, &m->d_losort, m->h_losort.data(), nFaces));
    LDU_TRY(upload(ctx, &m->d_lowerCol, lowerCol.data(), nFaces));
    {
        bool fits = nCells < (1 << 26);
        for (int c = 0; c < nCells && fits; c++) fits = m->h_ownerStart[c + 1] - m->h_ownerStart[c] <= 32;
        if (fits && nFaces) {
            std::vector<int> packed(nFaces);
            for (int k = 0; k < nFaces; k++) {
                const int f = m->h_losort[k], l = m->h_l[f];
                packed[k] = (l << 5) | (f - m->h_ownerStart[l]);
            }
            LDU_TRY(upload(ctx, &m->d_lowerPacked, packed.data(), nFaces, 8));
            // row blocks for the staged kernel
            const int nb = (nCells + kRowBlock - 1) / kRowBlock;
            std::vector<int> desc(4 * (size_t)nb);
            int fcap = 0, kcap = 0;
            for (int b = 0; b < nb; b++) {
                const int r0 = b * kRowBlock, r1 = std::min(nCells, r0 + kRowBlock);
                const int f0a = m->h_ownerStart[r0] & ~3, k0a = m->h_losortStart[r0] & ~3;
                const int nFa = (m->h_ownerStart[r1] - f0a + 3) & ~3, nKa = (m->h_losortStart[r1] - k0a + 3) & ~3;
                desc[4 * b] = f0a;
                desc[4 * b + 1] = nFa;
                desc[4 * b + 2] = k0a;
                desc[4 * b + 3] = nKa;
                fcap = std::max(fcap, nFa);
                kcap = std::max(kcap, nKa);
            }
            int* d = nullptr;
            LDU_TRY(upload(ctx, &d, desc.data(), desc.size(), 8));
            m->d_rowBlocks = d;
            m->nRowBlocks = nb;
            m->rowFaceCap = fcap;
            m->rowLowerCap = kcap;
        }
    }
    LDU_CUDA(cudaMalloc((void**)&m->d_diag, std::max(nCells, 1) * sizeof(double)));
    LDU_CUDA(cudaMalloc((void**)&m->d_upper, ((size_t)nFaces + 8) * sizeof(double)));
    LDU_CUDA(cudaMemsetAsync(m->d_upper + nFaces, 0, 8 * sizeof(double), ctx->stream));
    m->d_lower = m->d_upper;

    // interfaces, concatenated interface-major
    std::vector<int> cells;
    for (int i = 0; i < nInterfaces; i++) {
        Interface it;
        it.n = ifaceSizes[i];
        it.nbrRank = nbrRank ? nbrRank[i] : 0;
        it.nbrInterface = nbrInterface ? nbrInterface[i] : i;
        it.offset = (int)cells.size();
        for (int k = 0; k < it.n; k++) {
            const int c = faceCells[i][k];
            if (c < 0 || c >= nCells) {
                set_error("ldu_matrix_create: interface faceCells out of range");
                return LDU_EINVAL;
            }
            cells.push_back(c);
        }
        m->ifs.push_back(it);
    }
    m->nIfFaces = (int)cells.size();
    if (m->nIfFaces) {
        LDU_TRY(upload(ctx, &m->d_ifCells, cells.data(), cells.size()));
        LDU_CUDA(cudaMalloc((void**)&m->d_bou, cells.size() * sizeof(double)));
        LDU_CUDA(cudaMalloc((void**)&m->d_int, cells.size() * sizeof(double)));
        LDU_CUDA(cudaMalloc((void**)&m->d_recv, cells.size() * sizeof(double)));
        // boundary rows: entries of one cell in (interface, face) order == the
        // order updateMatrixInterfaces applies them in (patch loop, then face loop)
        std::vector<int> order(cells.size());
        for (size_t k = 0; k < order.size(); k++) order[k] = (int)k;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cells[a] < cells[b]; });
        std::vector<int> rowCell, rowStart;
        for (size_t k = 0; k < order.size(); k++) {
            if (k == 0 || cells[order[k]] != cells[order[k - 1]]) {
                rowCell.push_back(cells[order[k]]);
                rowStart.push_back((int)k);
            }
        }
        rowStart.push_back((int)order.size());
        m->nBRows = (int)rowCell.size();
        LDU_TRY(upload(ctx, &m->d_bRowCell, rowCell.data(), rowCell.size()));
        LDU_TRY(upload(ctx, &m->d_bRowStart, rowStart.data(), rowStart.size()));
        LDU_TRY(upload(ctx, &m->d_bEntry, order.data(), order.size()));
        m->ifBlockStart = rowCell.front();   // boundary rows are sorted by cell
    }
    LDU_TRY(ensure_scalars(m));
    LDU_CUDA(cudaStreamSynchronize(ctx->stream));
    guard.m = nullptr;
    *out = m;
    return LDU_OK;
}

static void free_schedule(Schedule& s)
{
    cudaFree(s.d_rows);
    cudaFree(s.d_levelStart);
    s = Schedule();
}

int ldu_matrix_destroy(ldu_matrix* m)
{
    if (!m) return LDU_OK;
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    gamg_free(m);
    cudaFree(m->d_l);
    cudaFree(m->d_u);
    cudaFree(m->d_ownerStart);
    cudaFree(m->d_losortStart);
    cudaFree(m->d_losort);
    cudaFree(m->d_lowerCol);
    cudaFree(m->d_lowerPacked);
    cudaFree(m->d_rowBlocks);
    cudaFree(m->d_diag);
    cudaFree(m->d_upper);
    if (m->ownLower) cudaFree(m->d_lower);
    cudaFree(m->d_ifCells);
    cudaFree(m->d_bou);
    cudaFree(m->d_int);
    cudaFree(m->d_recv);
    cudaFree(m->d_ifTable);
    cudaFree(m->d_bRowCell);
    cudaFree(m->d_bRowStart);
    cudaFree(m->d_bEntry);
    cudaFree(m->d_cellBRow);
    cudaFree(m->d_mcRows);
    cudaFree(m->d_ownerByNbrDesc);
    flow_free(m);
    stencil_free(m);
    stencil2_free(m);
    free_schedule(m->fwd);
    free_schedule(m->bwd);
    for (double* w : m->work) cudaFree(w);
    cudaFree(m->d_scalars);
    cudaFree(m->d_hist);
    delete m;
    return LDU_OK;
}

static int set_lower_storage(ldu_matrix* m, bool asym)
{
    if (asym && !m->ownLower) {
        LDU_CUDA(cudaMalloc((void**)&m->d_lower, ((size_t)m->nFaces + 8) * sizeof(double)));
        LDU_CUDA(cudaMemsetAsync(m->d_lower + m->nFaces, 0, 8 * sizeof(double), m->ctx->stream));
        m->ownLower = true;
    } else if (!asym && m->ownLower) {
        LDU_CUDA(cudaStreamSynchronize(m->ctx->stream));
        cudaFree(m->d_lower);
        m->d_lower = m->d_upper;
        m->ownLower = false;
    }
    if (!asym) m->d_lower = m->d_upper;
    m->symmetric = !asym;
    return LDU_OK;
}

int ldu_matrix_set_coeffs(ldu_matrix* m, const double* diag, const double* upper,
                          const double* lower, const double* const* bouCoeffs,
                          const double* const* intCoeffs)
{
    if (!m || !diag || (m->nFaces && !upper)) {
        set_error("ldu_matrix_set_coeffs: bad argument");
        return LDU_EINVAL;
    }
    cudaStream_t st = m->ctx->stream;
    LDU_CUDA(cudaSetDevice(m->ctx->device));
    LDU_TRY(set_lower_storage(m, lower != nullptr));
    LDU_TRY(copy_h2d(m->ctx, m->d_diag, diag, m->nCells * sizeof(double)));
    if (m->nFaces) {
        LDU_TRY(copy_h2d(m->ctx, m->d_upper, upper, m->nFaces * sizeof(double)));
        if (lower) LDU_TRY(copy_h2d(m->ctx, m->d_lower, lower, m->nFaces * sizeof(double)));
    }
    for (size_t i = 0; i < m->ifs.size(); i++) {
        const Interface& it = m->ifs[i];
        if (!it.n) continue;
        if (!bouCoeffs || !intCoeffs || !bouCoeffs[i] || !intCoeffs[i]) {
            set_error("ldu_matrix_set_coeffs: interface coefficients missing");
            return LDU_EINVAL;
        }
        LDU_CUDA(cudaMemcpyAsync(m->d_bou + it.offset, bouCoeffs[i], it.n * sizeof(double),
                                 cudaMemcpyHostToDevice, st));
        LDU_CUDA(cudaMemcpyAsync(m->d_int + it.offset, intCoeffs[i], it.n * sizeof(double),
                                 cudaMemcpyHostToDevice, st));
    }
    // the host arrays may be pageable and reused by the caller right away
    LDU_CUDA(cudaStreamSynchronize(st));
    m->diagonalOnly = (upper == nullptr && lower == nullptr);
    m->haveCoeffs = true;
    m->haveIfCoeffs = true;
    m->coefGen++;
    return LDU_OK;
}

int ldu_matrix_set_interface_coeffs(ldu_matrix* m, const double* const* bouCoeffs, const double* const* intCoeffs)
{
    if (!m) return LDU_EINVAL;
    LDU_CUDA(cudaSetDevice(m->ctx->device));
    cudaStream_t st = m->ctx->stream;
    for (size_t i = 0; i < m->ifs.size(); i++) {
        const Interface& it = m->ifs[i];
        if (!it.n) continue;
        if (!bouCoeffs || !intCoeffs || !bouCoeffs[i] || !intCoeffs[i]) {
            set_error("ldu_matrix_set_interface_coeffs: interface coefficients missing");
            return LDU_EINVAL;
        }
        LDU_CUDA(cudaMemcpyAsync(m->d_bou + it.offset, bouCoeffs[i], it.n * sizeof(double),
                                 cudaMemcpyHostToDevice, st));
        LDU_CUDA(cudaMemcpyAsync(m->d_int + it.offset, intCoeffs[i], it.n * sizeof(double),
                                 cudaMemcpyHostToDevice, st));
    }
    LDU_CUDA(cudaStreamSynchronize(st));
    m->haveIfCoeffs = true;
    m->coefGen++;
    return LDU_OK;
}

int ldu_matrix_set_coeffs_device(ldu_matrix* m, const double* d_diag, const double* d_upper,
                                 const double* d_lower)
{
    if (!m || !d_diag || (m->nFaces && !d_upper)) {
        set_error("ldu_matrix_set_coeffs_device: bad argument");
        return LDU_EINVAL;
    }
    if (m->nIfFaces > 0 && !m->haveIfCoeffs) {
        // the coupled rows would run on uninitialised boundary coefficients
        set_error("ldu_matrix_set_coeffs_device: the matrix has coupled patches; give their coefficients first "
                  "(ldu_matrix_set_interface_coeffs or ldu_matrix_set_coeffs)");
        return LDU_EINVAL;
    }
    LDU_CUDA(cudaSetDevice(m->ctx->device));
    cudaStream_t st = m->ctx->stream;
    LDU_TRY(set_lower_storage(m, d_lower != nullptr));
    LDU_CUDA(cudaMemcpyAsync(m->d_diag, d_diag, m->nCells * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (m->nFaces) {
        LDU_CUDA(cudaMemcpyAsync(m->d_upper, d_upper, m->nFaces * sizeof(double), cudaMemcpyDeviceToDevice, st));
        if (d_lower)
            LDU_CUDA(cudaMemcpyAsync(m->d_lower, d_lower, m->nFaces * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    m->coefGen++;
    m->diagonalOnly = (d_upper == nullptr && d_lower == nullptr);
    m->haveCoeffs = true;
    return LDU_OK;
}

// Colour-ordered renumbering (the role renumberMesh plays for bandwidth,
// here for dependency depth): greedy first-fit colouring of the cell graph in cell order, then
// cells sorted by (colour, old index).  In the new numbering a cell's lower neighbours all have
// smaller colours, so the lexicographic Gauss-Seidel / DIC sweeps of the reference have a
// dependency depth equal to the number of colours (2 on a hex box) instead of nx+ny+nz:
// the reference's own smoother IS a multi-colour smoother on such a mesh.  Host only.
int ldu_colour_order(int nCells, int nFaces, const int* lowerAddr, const int* upperAddr, int* newIndexOfOldCell,
                     int* nColours)
{
    if (nCells < 0 || nFaces < 0 || !newIndexOfOldCell || (nFaces && (!lowerAddr || !upperAddr))) {
        set_error("ldu_colour_order: bad argument");
        return LDU_EINVAL;
    }
    std::vector<int> colour;
    int nc = 0;
    LDU_TRY(ldu::greedy_colouring(nCells, nFaces, lowerAddr, upperAddr, colour, &nc));
    // stable counting sort by colour
    std::vector<int> first(nc + 1, 0);
    for (int c = 0; c < nCells; c++) first[colour[c] + 1]++;
    for (int q = 0; q < nc; q++) first[q + 1] += first[q];
    for (int c = 0; c < nCells; c++) newIndexOfOldCell[c] = first[colour[c]]++;
    if (nColours) *nColours = nc;
    return LDU_OK;
}

// Cuthill-McKee band compression as renumberMesh's default method applies it
// (renumber/renumberMethods/CuthillMcKeeRenumber -> meshes/bandCompression/bandCompression.C:42-148):
// breadth-first walk per connected component, each started from the lowest-numbered unvisited cell of
// minimal degree; a visited cell appends its unvisited neighbours to the queue.  The reference
// computes the degree-sorted order of those neighbours and then appends nbrs[i], not nbrs[order[i]]
// (bandCompression.C:133-139) -- the neighbours go in cell-cell order (ascending face index); restated
// as is, so that the numbering equals what renumberMesh produces.  Host only.
// oldCellOfNewCell[new] = old (the reference's "newOrder").
int ldu_band_compression(int nCells, int nFaces, const int* lowerAddr, const int* upperAddr,
                         int* oldCellOfNewCell)
{
    if (nCells < 0 || nFaces < 0 || !oldCellOfNewCell || (nFaces && (!lowerAddr || !upperAddr))) {
        set_error("ldu_band_compression: bad argument");
        return LDU_EINVAL;
    }
    // cell-cell addressing in face order (decompositionMethod::calcCellCells on the internal faces)
    std::vector<int> start(nCells + 1, 0);
    for (int f = 0; f < nFaces; f++) {
        const int l = lowerAddr[f], u = upperAddr[f];
        if (l < 0 || u < 0 || l >= nCells || u >= nCells || l == u) {
            set_error("ldu_band_compression: addressing out of range");
            return LDU_EINVAL;
        }
        start[l + 1]++;
        start[u + 1]++;
    }
    for (int c = 0; c < nCells; c++) start[c + 1] += start[c];
    std::vector<int> adj(start[nCells]), fill(start.begin(), start.end() - 1);
    for (int f = 0; f < nFaces; f++) {
        adj[fill[lowerAddr[f]]++] = upperAddr[f];
        adj[fill[upperAddr[f]]++] = lowerAddr[f];
    }
    std::vector<char> visited(nCells, 0);
    std::vector<int> queue;
    queue.reserve((size_t)nCells + adj.size());
    int placed = 0;
    for (;;) {
        int current = -1, minDegree = 0x7fffffff;
        for (int c = 0; c < nCells; c++) {
            if (!visited[c] && start[c + 1] - start[c] < minDegree) {
                minDegree = start[c + 1] - start[c];
                current = c;
            }
        }
        if (current < 0) break;
        queue.clear();
        queue.push_back(current);
        for (size_t head = 0; head < queue.size(); head++) {
            const int c = queue[head];
            if (visited[c]) continue;
            visited[c] = 1;
            oldCellOfNewCell[placed++] = c;
            for (int k = start[c]; k < start[c + 1]; k++)
                if (!visited[adj[k]]) queue.push_back(adj[k]);
        }
    }
    return LDU_OK;
}

int ldu_matrix_set_face_weights(ldu_matrix* m, const double* weights)
{
    if (!m || (m->nFaces && !weights)) return LDU_EINVAL;
    m->h_faceWeights.assign(weights, weights + m->nFaces);
    if (!m->externalHierarchy) m->hierarchyValid = false;
    return LDU_OK;
}

}  // extern "C"
