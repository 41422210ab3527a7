// Host side of the multi-GPU exchange + the halo put / receive kernels.
//
// One process per GPU, one mesh region per GPU (the reference's decomposePar /
// MPI model).  The only traffic between regions is
//   * processor-patch halos: psi[faceCells] of every coupled patch, once per
//     Amul / Tmul / residual / Gauss-Seidel sweep and per GAMG level
//     (lduMatrixUpdateMatrixInterfaces.C:30-266, processorFvPatchScalarField.C:36-144)
//   * 1-3 double sums per reduction (comm.cuh)
// Both are written by the producing GPU directly into the consumer's exchange
// window (CUDA IPC mapping, NVLink/NVSwitch P2P stores) from inside the kernels
// that produce them; the consumer spins on an epoch flag in its own HBM.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "comm.cuh"
#include "reduce.cuh"

namespace ldu {

CommDev comm_dev(const ldu_context* ctx)
{
    CommDev c;
    c.rank = ctx->comm.rank;
    c.nRanks = ctx->comm.connected ? ctx->comm.nRanks : 1;
    c.peer = ctx->comm.d_peer;
    c.maxInterfaces = ctx->comm.maxInterfaces;
    c.slotStride = ctx->comm.slotStride;
    c.timeoutCycles = 20000000000ll;  // ~10 s: fail loudly instead of hanging the GPU
    static const int llRed = (getenv("LDU_RED_LL") && getenv("LDU_RED_LL")[0] == '0') ? 0 : 1;
    c.llRed = llRed;
    return c;
}

// psi[faceCells] of every interface -> the neighbour's window, then publish the
// epoch to every neighbour (last block).
__global__ void __launch_bounds__(kBlock) halo_put_kernel(CommDev c, const IfaceDev* __restrict__ ifs,
                                                           int nIfs, const int* __restrict__ ifCells,
                                                           const double* __restrict__ psi,
                                                           const SolverScalars* __restrict__ guard)
{
    if (guard && guard->done) return;
    WindowHeader* me = win_hdr(c, c.rank);
    const IfaceDev it = ifs[blockIdx.y];
    const int par = (int)((me->haloSent[it.nbrRank] + 1ull) & 1ull);     // the pair's next exchange
    double* dst = win_halo(c, it.nbrRank, par, it.nbrInterface);
    for (int i = blockIdx.x * kBlock + threadIdx.x; i < it.n; i += gridDim.x * kBlock)
        dst[i] = psi[ifCells[it.offset + i]];
    __threadfence_system();
    __shared__ bool isLast;
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(&me->haloTicket, 1u);
        isLast = (t == gridDim.x * gridDim.y - 1);
    }
    __syncthreads();
    if (!isLast) return;
    if (threadIdx.x == 0) {
        __threadfence_system();
        me->haloTicket = 0u;
        for (int k = 0; k < nIfs; k++) {
            // one counter and one flag per neighbour rank, however many interfaces lead to it
            const int r = ifs[k].nbrRank;
            bool first = true;
            for (int j = 0; j < k; j++) first = first && ifs[j].nbrRank != r;
            if (!first) continue;
            const unsigned long long epoch = me->haloSent[r] + 1ull;
            me->haloSent[r] = epoch;
            st_release_sys(&win_hdr(c, r)->haloSeq[(int)(epoch & 1ull)][c.rank], epoch);
        }
    }
}

// wait for every neighbour's halo of the current epoch, then gather the window
// slots into the matrix's concatenated receive buffer
__global__ void __launch_bounds__(kBlock) halo_recv_kernel(CommDev c, const IfaceDev* __restrict__ ifs,
                                                            int nIfs, double* __restrict__ recv,
                                                            SolverScalars* __restrict__ S, bool guarded)
{
    if (guarded && S->done) return;
    WindowHeader* me = win_hdr(c, c.rank);
    __shared__ bool ok;
    if (threadIdx.x == 0) {
        ok = true;
        for (int k = 0; k < nIfs && ok; k++) {
            // the neighbour's message of the exchange this rank has just sent its own for
            const unsigned long long epoch = me->haloSent[ifs[k].nbrRank];
            ok = wait_epoch(&me->haloSeq[(int)(epoch & 1ull)][ifs[k].nbrRank], epoch, c.timeoutCycles);
        }
        if (!ok) {
            S->commError = 1;
            S->done = 1;
        }
    }
    __syncthreads();
    if (!ok) return;
    const IfaceDev it = ifs[blockIdx.y];
    const int par = (int)(me->haloSent[it.nbrRank] & 1ull);
    const double* src = win_halo(c, c.rank, par, blockIdx.y);
    for (int i = blockIdx.x * kBlock + threadIdx.x; i < it.n; i += gridDim.x * kBlock)
        recv[it.offset + i] = ld_volatile_f64(src + i);
}

static int ensure_iface_table(ldu_matrix* m, IfaceDev** out)
{
    // cached in a work slot of its own (tiny)
    static_assert(sizeof(IfaceDev) == 16, "IfaceDev layout");
    if (!m->d_ifTable) {
        std::vector<IfaceDev> h(m->ifs.size());
        for (size_t i = 0; i < h.size(); i++)
            h[i] = IfaceDev{m->ifs[i].offset, m->ifs[i].n, m->ifs[i].nbrRank, m->ifs[i].nbrInterface};
        LDU_CUDA(cudaMalloc((void**)&m->d_ifTable, std::max<size_t>(h.size(), 1) * sizeof(IfaceDev)));
        LDU_CUDA(cudaMemcpy(m->d_ifTable, h.data(), h.size() * sizeof(IfaceDev), cudaMemcpyHostToDevice));
    }
    *out = reinterpret_cast<IfaceDev*>(m->d_ifTable);
    return LDU_OK;
}

static int halo_grid(ldu_matrix* m, dim3& grid, IfaceDev** tab)
{
    ldu_context* ctx = m->ctx;
    int maxN = 0;
    bool allSelf = true;
    for (const Interface& it : m->ifs) {
        maxN = std::max(maxN, it.n);
        allSelf = allSelf && it.nbrRank == 0;
    }
    if (!ctx->comm.connected || (ctx->comm.selfOnly && allSelf
                                 && ((int)m->ifs.size() > ctx->comm.maxInterfaces || maxN > ctx->comm.slotStride))) {
        // a single process whose coupled patches all point back into the region itself (cyclic
        // pairs): it is its own only peer, the window is made (or re-made larger) on demand
        if (!allSelf || (ctx->comm.connected && !ctx->comm.selfOnly)) {
            set_error("matrix has coupled interfaces but the context has no peers (ldu_comm_connect)");
            return LDU_ECOMM;
        }
        if (ctx->comm.connected) {
            LDU_CUDA(cudaStreamSynchronize(ctx->stream));
            cudaFree(ctx->comm.window);
            cudaFree(ctx->comm.d_peer);
            ctx->comm.window = nullptr;
            ctx->comm.d_peer = nullptr;
            ctx->comm.connected = false;
        }
        unsigned char handle[LDU_COMM_HANDLE_BYTES];
        LDU_TRY(ldu_comm_window_create(ctx, 0, 1, (int)m->ifs.size(), maxN, handle));
        LDU_TRY(ldu_comm_connect(ctx, handle));
        ctx->comm.selfOnly = true;
    }
    if (ctx->comm.selfOnly && !allSelf) {
        set_error("matrix has interfaces to other ranks but the context has no peers (ldu_comm_connect)");
        return LDU_ECOMM;
    }
    if ((int)m->ifs.size() > ctx->comm.maxInterfaces || maxN > ctx->comm.slotStride) {
        set_error("exchange window too small for this matrix's interfaces");
        return LDU_ECOMM;
    }
    // a wrong neighbour rank / interface index would be an out-of-bounds store into a peer's HBM
    for (const Interface& it : m->ifs) {
        if (it.nbrRank < 0 || it.nbrRank >= ctx->comm.nRanks || it.nbrInterface < 0
            || it.nbrInterface >= ctx->comm.maxInterfaces) {
            set_error("interface names a neighbour rank / interface outside the exchange window "
                      "(nbrRank < nRanks and nbrInterface < maxInterfaces of ldu_comm_window_create)");
            return LDU_EINVAL;
        }
    }
    LDU_TRY(ensure_iface_table(m, tab));
    grid = dim3(std::max(1, std::min(16, (maxN + kBlock - 1) / kBlock)), (unsigned)m->ifs.size());
    return LDU_OK;
}

// initMatrixInterfaces (lduMatrixUpdateMatrixInterfaces.C:30-93): start the halo
// sends.  Issued BEFORE the interior row kernel so the NVLink transfer overlaps it,
// exactly the reference's init/update split (lduMatrixATmul.C:57-64,82-89).
int comm_halo_put(ldu_matrix* m, const double* d_psi, bool guarded)
{
    if (!m->nIfFaces) return LDU_OK;
    dim3 grid;
    IfaceDev* tab;
    LDU_TRY(halo_grid(m, grid, &tab));
    halo_put_kernel<<<grid, kBlock, 0, m->ctx->stream>>>(comm_dev(m->ctx), tab, (int)m->ifs.size(), m->d_ifCells,
                                                         d_psi, guarded ? m->d_scalars : nullptr);
    count_launch();
    LDU_CUDA(cudaGetLastError());
    return LDU_OK;
}

// updateMatrixInterfaces, first half: wait for the neighbours' halos
int comm_halo_recv(ldu_matrix* m, bool guarded)
{
    if (!m->nIfFaces) return LDU_OK;
    dim3 grid;
    IfaceDev* tab;
    LDU_TRY(halo_grid(m, grid, &tab));
    halo_recv_kernel<<<grid, kBlock, 0, m->ctx->stream>>>(comm_dev(m->ctx), tab, (int)m->ifs.size(), m->d_recv,
                                                          m->d_scalars, guarded);
    count_launch();
    LDU_CUDA(cudaGetLastError());
    return LDU_OK;
}

// the validated device table of a matrix's interfaces, for kernels that exchange halos themselves (gamg.cu)
int comm_halo_table(ldu_matrix* m, const IfaceDev** tab)
{
    *tab = nullptr;
    if (!m->nIfFaces) return LDU_OK;
    dim3 grid;
    IfaceDev* t = nullptr;
    LDU_TRY(halo_grid(m, grid, &t));
    *tab = t;
    return LDU_OK;
}

int comm_halo_exchange(ldu_matrix* m, const double* d_psi, bool guarded)
{
    LDU_TRY(comm_halo_put(m, d_psi, guarded));
    return comm_halo_recv(m, guarded);
}

__global__ void allreduce_kernel(CommDev c, double* vals, int n, SolverScalars* S)
{
    double v[kRedSlots];
    for (int k = 0; k < kRedSlots; k++) v[k] = k < n ? vals[k] : 0.0;
    comm_allreduce_dev<kRedSlots>(c, v, S);
    for (int k = 0; k < n; k++) vals[k] = v[k];
}

int comm_allreduce(ldu_context* ctx, double* d_vals, int n)
{
    if (!ctx->comm.connected || ctx->comm.nRanks == 1) return LDU_OK;
    if (n > kRedSlots) return LDU_EINVAL;
    allreduce_kernel<<<1, 1, 0, ctx->stream>>>(comm_dev(ctx), d_vals, n, nullptr);
    count_launch();
    LDU_CUDA(cudaGetLastError());
    return LDU_OK;
}

}  // namespace ldu

using namespace ldu;

extern "C" {

int ldu_comm_window_create(ldu_context* ctx, int rank, int nRanks, int maxInterfaces,
                           long long maxInterfaceFaces, unsigned char* handleOut)
{
    if (!ctx || !handleOut || nRanks < 1 || nRanks > kMaxRanks || rank < 0 || rank >= nRanks) {
        set_error("ldu_comm_window_create: bad argument");
        return LDU_EINVAL;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) <= LDU_COMM_HANDLE_BYTES, "handle size");
    LDU_CUDA(cudaSetDevice(ctx->device));
    Comm& cm = ctx->comm;
    if (cm.window) {   // a window the library made for cyclic-only matrices, or an earlier call: replace it
        LDU_CUDA(cudaStreamSynchronize(ctx->stream));
        if (cm.connected)      // the peers' windows mapped by the earlier ldu_comm_connect
            for (int r = 0; r < cm.nRanks; r++)
                if (r != cm.rank && cm.peer[r]) {
                    cudaIpcCloseMemHandle(cm.peer[r]);
                    cm.peer[r] = nullptr;
                }
        cudaFree(cm.window);
        cudaFree(cm.d_peer);
        cm.window = nullptr;
        cm.d_peer = nullptr;
        cm.connected = false;
    }
    cm.selfOnly = false;
    cm.rank = rank;
    cm.nRanks = nRanks;
    cm.maxInterfaces = std::max(maxInterfaces, 1);
    cm.slotStride = std::max<long long>(maxInterfaceFaces, 1);
    cm.windowBytes = sizeof(WindowHeader) + (size_t)2 * cm.maxInterfaces * cm.slotStride * sizeof(double);
    LDU_CUDA(cudaMalloc((void**)&cm.window, cm.windowBytes));
    LDU_CUDA(cudaMemset(cm.window, 0, cm.windowBytes));
    memset(handleOut, 0, LDU_COMM_HANDLE_BYTES);
    if (nRanks > 1) {
        cudaIpcMemHandle_t h;
        LDU_CUDA(cudaIpcGetMemHandle(&h, cm.window));
        memcpy(handleOut, &h, sizeof(h));
    }
    return LDU_OK;
}

int ldu_comm_connect(ldu_context* ctx, const unsigned char* allHandles)
{
    if (!ctx || !ctx->comm.window) {
        set_error("ldu_comm_connect: create the window first");
        return LDU_EINVAL;
    }
    Comm& cm = ctx->comm;
    LDU_CUDA(cudaSetDevice(ctx->device));
    for (int r = 0; r < cm.nRanks; r++) {
        if (r == cm.rank) {
            cm.peer[r] = cm.window;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, allHandles + (size_t)r * LDU_COMM_HANDLE_BYTES, sizeof(h));
        void* p = nullptr;
        LDU_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        cm.peer[r] = (unsigned char*)p;
    }
    LDU_CUDA(cudaMalloc((void**)&cm.d_peer, kMaxRanks * sizeof(unsigned char*)));
    LDU_CUDA(cudaMemcpy(cm.d_peer, cm.peer, kMaxRanks * sizeof(unsigned char*), cudaMemcpyHostToDevice));
    cm.connected = true;
    return LDU_OK;
}

}  // extern "C"
