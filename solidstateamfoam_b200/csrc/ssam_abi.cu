// C-ABI of the B200-native constitutive + heat-source update (include/ssam_gpu.h).
// Host-side orchestration only: every number is produced by the kernels in kernels.cuh.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "kernels.cuh"
#include <cuda.h>

#include "nccl_dyn.h"
#include "ssam_gpu.h"

using namespace ssam;

struct ssam_ctx {
    int device = 0;
    cudaStream_t ownStream = nullptr, stream = nullptr;
    bool blocking = true;
    std::string err;
    int64_t launches = 0;

    int nCells = 0, nIF = 0, nFaces = 0, nBnd = 0, nTot = 0, nPatches = 0;
    bool haveMesh = false;
    std::vector<int> patchStart, patchSize, patchKind, patchNbrRank, patchUbc;
    std::vector<unsigned char> bndFlagsHost;

    FaceGeo* geo = nullptr;
    int *owner = nullptr, *neighbour = nullptr, *ownerStart = nullptr, *extraStart = nullptr;
    int2* extra = nullptr;
    int nExtra = 0;
    int4* tileDesc = nullptr;
    int nTiles = 0;
    int *tileListInterior = nullptr, *tileListCoupled = nullptr;
    int nTilesInterior = 0, nTilesCoupled = 0;
    int* tileOrder = nullptr;  // L2-friendly traversal order of the tiles (nullptr = natural order)
    int4 *descInterior = nullptr, *descCoupled = nullptr, *descOrder = nullptr;  // tileDesc in list order
    int interiorRange0 = -1;  // >= 0: the interior tiles are the contiguous range [interiorRange0, +nTilesInterior)
    int* tileIota = nullptr;  // identity tile list (chunked launches of the host-staged pipeline)
    cudaStream_t h2dStream = nullptr, d2hStream = nullptr;
    cudaStream_t commStream = nullptr;
    cudaEvent_t evMainToComm = nullptr, evCommToMain = nullptr;
    cudaEvent_t evHaloIn = nullptr, evCoupledDone = nullptr;  // peer schedule: inputs unpacked / coupled tiles finished
    int pfDist = 148;    // L2 prefetch distance of the tiled kernel, in tiles (measured optimum, profiles/)
    int bandwidthTiles = 0;  // ceil(max(neighbour - owner) / TILE)
    // 97th percentile of the tile distance of the internal faces: where the HEAVY forward neighbours sit (a tile order
    // may put a few light neighbours very far away; prefetching towards those would only waste DRAM traffic)
    int pfBandwidthTiles = 0;
    int hostChunks = -1;  // >= 0 overrides the chunk count of the host-staged pipeline (0/1 = unpipelined)
    int nTilesOversize = 0;
    // 0 = one thread per cell, direct loads; 1 (default) = tiled, bulk-async staged, one CTA per tile; 2 = persistent
    // self-pipelined tiles (round-2 experiment: slower than 1 on every workload, profiles/r2_k_cells_pipe.md)
    int cellKernel = 1;
    int pipeGridCap = 0;  // > 0: cap on the persistent kernel's grid (experiments)
    int* pipeSched = nullptr;  // 8 x {next position, CTAs done}: counters of the persistent kernel's dynamic scheduler
    int pipeSchedNext = 0;     // launches cycle through the 8 pairs (two launches of one update may overlap)
    // cell layout: 0 = the caller's numbering, 1 = tile-compact when it pays (default), 2 = always tile-compact
    int layoutMode = 1;
    int* perm = nullptr;      // device position i holds the caller's cell perm[i] (nullptr = identity)
    int* faceKey = nullptr;   // internal faces: index of the face in the caller's mesh (permuted layout only)
    std::vector<int> permHost;
    int* invPerm = nullptr;             // caller cell -> device position (permuted layout only)
    std::vector<int> tileMaxCallerCell;  // per tile, for the host-staged pipeline (built on first use)
    std::vector<int> pipeWaitIn, pipeOutWait;  // chunk dependencies of that pipeline, cached per chunk count
    int pipeChunks = 0;
    // launch order of the host-staged pipeline on a permuted layout: tiles sorted by their first caller cell, so that
    // upload chunk -> compute chunk -> download chunk stay monotone whatever order the tiles have in memory
    int* pipeList = nullptr;
    int4* descPipe = nullptr;
    std::vector<int> pipeListHost, pipePosOfTile;
    int *origOwner = nullptr, *origNeighbour = nullptr;  // the caller's owner / neighbour (permuted layout only)
    ssam_ctx* refAddr = nullptr;  // addressing tables in the caller's numbering, built on demand
    double remotePerCellCaller = 0.0, remotePerCellCompact = 0.0;
    double *V = nullptr, *magSf_b = nullptr, *deltaCoeffs_b = nullptr;
    double* Cpl[3] = {nullptr, nullptr, nullptr};  // cell centres, then boundary-face centres
    unsigned char* bndFlags = nullptr;
    double* gradB[9] = {nullptr};
    double* faceField[SSAM_FF_COUNT] = {nullptr};  // surfaceScalarFields (nFaces values each), lazily allocated
    bool faceFieldSet[SSAM_FF_COUNT] = {false};
    double* minmaxScratch = nullptr;
    double* procNbrC[3] = {nullptr, nullptr, nullptr};  // [nBnd] neighbour-rank cell centre behind each processor face
    bool procNbrCSet = false;
    double sumV = 0.0;        // sequential sum of the cell volumes (deltaN_)

    double* plane[SSAM_F_COUNT][9] = {{nullptr}};
    bool isSet[SSAM_F_COUNT] = {false};

    double* stage = nullptr;
    size_t stageCap = 0;

    int fricPatchCached = -1;
    bool fusedPrepared = false;  // ssam_prepare_fused ran on this mesh (in-process rank groups require it)
    // boundary-face centres and areas as given by the caller: the friction-patch sums of an in-process rank group are
    // formed on the host (pure mesh geometry; no device synchronisation inside a collective update)
    std::vector<double> hostCfB, hostMagSfB;
    double fricHost[4] = {0, 0, 0, 0};
    double* fricSums = nullptr;  // device: [0..3] local sums, [4..7] cross-rank sums
    int qfricZeroedFor = -2;
    bool nu2Filled = false;
    double nu2FilledWith = 0.0;

    // halo
    int nProcPatches = 0, nProcFaces = 0;
    std::vector<int> haloPatch;    // patch index of each processor patch
    std::vector<int> haloOffsets;  // offsets into the send-cell list
    int* haloSendCells = nullptr;
    double* sendBuf = nullptr;
    size_t sendCap = 0;
    ncclComm_t comm = nullptr;
    int rank = 0, nRanks = 1;
    // in-process rank group (ssam_comm_init_local): every rank is a context of THIS process; halo slots are linked by
    // plain device pointers instead of IPC handles, the small reductions go through the host
    std::vector<ssam_ctx*> localPeers;
    // halo exchange over peer memory (IPC-mapped receive slots); NCCL stays the fallback
    struct P2P {
        int state = 0;  // 0 = not tried, 1 = on, -1 = unavailable (NCCL path)
        double* recvBuf = nullptr;  // [256 B of flag words][2 slots of slotDoubles]
        size_t slotDoubles = 0;
        unsigned* done = nullptr;
        int* patchB0 = nullptr;
        void* peerMap[32] = {nullptr};      // IPC mappings this context opened (one per neighbour rank)
        double* peerBase[32] = {nullptr};   // per processor patch: the neighbour's receive buffer
        int peerOff[32] = {0};
        size_t peerSlotDoubles[32] = {0};   // ... and the size of one of ITS two slots (ranks differ in processor faces)
        unsigned long long seq = 0;
    } p2p;

    size_t haloBytesLastUpdate = 0;  // bytes this rank sent to its neighbours during the last fused update
    // optional per-kernel timing (ssam_profile_enable)
    bool profile = false;
    int64_t profUpdates = 0;  // fused updates covered by profEvents (an update may contribute several event pairs)
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> profEvents;

    // lazily built integer tables
    int *cellFacesStart = nullptr, *cellFaces = nullptr, *faceColour = nullptr;
};

namespace {

thread_local std::string g_globalErr;

int ncompOf(int f) {
    return (f == SSAM_F_U || f == SSAM_F_GRADALPHA || f == SSAM_F_DIVDEVRHOREFF) ? 3 : (f == SSAM_F_GRADU ? 9 : 1);
}

ssam_status fail(ssam_ctx* c, ssam_status st, const std::string& msg) {
    if (c)
        c->err = msg;
    else
        g_globalErr = msg;
    return st;
}

#define CU(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(ctx, SSAM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));       \
    } while (0)
#define CHECK(cond, st, msg) \
    do {                     \
        if (!(cond)) return fail(ctx, st, msg); \
    } while (0)
#define TRY(call)                 \
    do {                          \
        ssam_status s_ = (call);  \
        if (s_ != SSAM_OK) return s_; \
    } while (0)
#define NC(call)                                                                                              \
    do {                                                                                                      \
        ncclResult_t r_ = (call);                                                                             \
        if (r_ != ncclSuccess)                                                                                \
            return fail(ctx, SSAM_ERR_COMM, std::string(#call) + ": " + nccl().GetErrorString(r_));          \
    } while (0)

inline int gridFor(long long n, int block) { return int((n + block - 1) / block); }

// launch + count + error check
#define LAUNCH(kernel, grid, block, ...)                                                  \
    do {                                                                                  \
        if ((grid) > 0) {                                                                 \
            kernel<<<(grid), (block), 0, ctx->stream>>>(__VA_ARGS__);                     \
            ctx->launches++;                                                              \
            CU(cudaGetLastError());                                                       \
        }                                                                                 \
    } while (0)

ssam_status finish(ssam_ctx* ctx) {
    if (ctx->blocking) CU(cudaStreamSynchronize(ctx->stream));
    return SSAM_OK;
}

template <class T>
ssam_status devAlloc(ssam_ctx* ctx, T** p, size_t n) {
    *p = nullptr;
    if (n == 0) n = 1;
    CU(cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T)));
    return SSAM_OK;
}
template <class T>
void devFree(T*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}

ssam_status ensureStage(ssam_ctx* ctx, size_t nDoubles) {
    if (ctx->stageCap >= nDoubles) return SSAM_OK;
    CU(cudaStreamSynchronize(ctx->stream));
    devFree(ctx->stage);
    TRY(devAlloc(ctx, &ctx->stage, nDoubles));
    ctx->stageCap = nDoubles;
    return SSAM_OK;
}

ssam_status ensureField(ssam_ctx* ctx, int f) {
    CHECK(ctx->haveMesh, SSAM_ERR_INVALID, "mesh not set");
    CHECK(f >= 0 && f < SSAM_F_COUNT, SSAM_ERR_INVALID, "bad field id");
    const int nc = ncompOf(f);
    for (int k = 0; k < nc; ++k) {
        if (ctx->plane[f][k]) continue;
        TRY(devAlloc(ctx, &ctx->plane[f][k], size_t(ctx->nTot) + 2));
        CU(cudaMemsetAsync(ctx->plane[f][k], 0, (size_t(ctx->nTot) + 2) * sizeof(double), ctx->stream));
    }
    return SSAM_OK;
}

ssam_status ensureFaceField(ssam_ctx* ctx, int ff) {
    CHECK(ctx->haveMesh, SSAM_ERR_INVALID, "mesh not set");
    if (!ctx->faceField[ff]) TRY(devAlloc(ctx, &ctx->faceField[ff], size_t(ctx->nFaces)));
    return SSAM_OK;
}

void p2pFree(ssam_ctx* c) {
    for (void*& b : c->p2p.peerMap) {
        if (b) cudaIpcCloseMemHandle(b);
        b = nullptr;
    }
    for (double*& b : c->p2p.peerBase) b = nullptr;
    devFree(c->p2p.recvBuf);
    devFree(c->p2p.done);
    devFree(c->p2p.patchB0);
    c->p2p.state = 0;
    c->p2p.seq = 0;
}

void freeMesh(ssam_ctx* c) {
    if (c->refAddr) {
        freeMesh(c->refAddr);
        delete c->refAddr;
        c->refAddr = nullptr;
    }
    devFree(c->geo);
    devFree(c->owner);
    devFree(c->neighbour);
    devFree(c->ownerStart);
    devFree(c->extraStart);
    devFree(c->extra);
    devFree(c->tileDesc);
    devFree(c->tileIota);
    devFree(c->pipeSched);
    devFree(c->tileOrder);
    devFree(c->descInterior);
    devFree(c->descCoupled);
    devFree(c->descOrder);
    devFree(c->perm);
    devFree(c->invPerm);
    c->tileMaxCallerCell.clear();
    c->pipeChunks = 0;
    devFree(c->pipeList);
    devFree(c->descPipe);
    c->pipeListHost.clear();
    c->pipePosOfTile.clear();
    devFree(c->faceKey);
    devFree(c->origOwner);
    devFree(c->origNeighbour);
    c->permHost.clear();
    devFree(c->tileListInterior);
    devFree(c->tileListCoupled);
    devFree(c->V);
    for (int k = 0; k < SSAM_FF_COUNT; ++k) {
        devFree(c->faceField[k]);
        c->faceFieldSet[k] = false;
    }
    devFree(c->minmaxScratch);
    for (auto& q : c->procNbrC) devFree(q);
    c->procNbrCSet = false;
    devFree(c->magSf_b);
    devFree(c->deltaCoeffs_b);
    for (auto& p : c->Cpl) devFree(p);
    devFree(c->bndFlags);
    for (auto& p : c->gradB) devFree(p);
    for (int f = 0; f < SSAM_F_COUNT; ++f) {
        for (auto& p : c->plane[f]) devFree(p);
        c->isSet[f] = false;
    }
    p2pFree(c);
    // in-process rank group: the neighbours' links into this context's (now freed) receive slots are re-made at their
    // next exchange
    for (ssam_ctx* q : c->localPeers)
        if (q != c) q->p2p.state = 0;
    devFree(c->haloSendCells);
    devFree(c->sendBuf);
    c->sendCap = 0;
    devFree(c->cellFacesStart);
    devFree(c->cellFaces);
    devFree(c->faceColour);
    c->fricPatchCached = -1;
    c->fusedPrepared = false;
    c->hostCfB.clear();
    c->hostMagSfB.clear();
    c->qfricZeroedFor = -2;
    c->nu2Filled = false;
    c->haveMesh = false;
}

// in-place exclusive scan of n ints on the device
ssam_status exclusiveScan(ssam_ctx* ctx, int* data, int n) {
    const int tiles = gridFor(n, SCAN_TILE);
    if (tiles <= 1) {
        LAUNCH(k_scan_tile, 1, SCAN_THREADS, data, n, (int*)nullptr);
        return SSAM_OK;
    }
    int* sums = nullptr;
    TRY(devAlloc(ctx, &sums, size_t(tiles)));
    LAUNCH(k_scan_tile, tiles, SCAN_THREADS, data, n, sums);
    ssam_status st = exclusiveScan(ctx, sums, tiles);
    if (st == SSAM_OK) {
        k_scan_add<<<gridFor(n, 256), 256, 0, ctx->stream>>>(data, n, sums);
        ctx->launches++;
    }
    cudaStreamSynchronize(ctx->stream);
    cudaFree(sums);
    return st;
}

void rebuildBndFlags(ssam_ctx* c) {
    c->bndFlagsHost.assign(size_t(std::max(c->nBnd, 1)), 0);
    for (int p = 0; p < c->nPatches; ++p) {
        unsigned char fl = 0;
        if (c->patchKind[p] == SSAM_PATCH_PROCESSOR)
            fl |= BND_COUPLED;
        else if (c->patchUbc[p] == SSAM_BC_ZERO_GRADIENT)
            fl |= BND_ZERO_GRADIENT;
        for (int f = c->patchStart[p]; f < c->patchStart[p] + c->patchSize[p]; ++f) c->bndFlagsHost[f - c->nIF] = fl;
    }
}

ssam_status buildAddressing(ssam_ctx* ctx) {
    const int nc = ctx->nCells, nIF = ctx->nIF, nF = ctx->nFaces;
    TRY(devAlloc(ctx, &ctx->ownerStart, size_t(nc) + 8));
    TRY(devAlloc(ctx, &ctx->extraStart, size_t(nc) + 8));
    int* errFlag = nullptr;
    TRY(devAlloc(ctx, &errFlag, 1));
    CU(cudaMemsetAsync(errFlag, 0, sizeof(int), ctx->stream));
    LAUNCH(k_owner_start, gridFor(nIF + 1, 256), 256, ctx->owner, nIF, nc, ctx->ownerStart, errFlag);
    int herr = 0;
    CU(cudaMemcpyAsync(&herr, errFlag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(errFlag);
    CHECK(herr == 0, SSAM_ERR_INVALID,
          "owner[] of the internal faces is not ascending: the mesh is not in upper-triangular order");
    CU(cudaMemsetAsync(ctx->extraStart, 0, (size_t(nc) + 1) * sizeof(int), ctx->stream));
    LAUNCH(k_extra_count, gridFor(nF, 256), 256, ctx->owner, ctx->neighbour, nIF, nF, ctx->extraStart);
    TRY(exclusiveScan(ctx, ctx->extraStart, nc + 1));
    ctx->nExtra = nIF + ctx->nBnd;
    TRY(devAlloc(ctx, &ctx->extra, size_t(ctx->nExtra) + 4));
    int* cursor = nullptr;
    TRY(devAlloc(ctx, &cursor, size_t(nc) + 1));
    CU(cudaMemcpyAsync(cursor, ctx->extraStart, (size_t(nc) + 1) * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
    LAUNCH(k_extra_fill, gridFor(nF, 256), 256, ctx->owner, ctx->neighbour, ctx->bndFlags, nIF, nF, cursor, ctx->extra);
    LAUNCH(k_extra_sort, gridFor(nc, 256), 256, ctx->extraStart, nc, ctx->extra, ctx->faceKey, nIF);
    {
        int* bw = nullptr;
        TRY(devAlloc(ctx, &bw, 1));
        CU(cudaMemsetAsync(bw, 0, sizeof(int), ctx->stream));
        LAUNCH(k_bandwidth, std::min(gridFor(nIF, 256), 148 * 8), 256, ctx->owner, ctx->neighbour, nIF, bw);
        int h = 0;
        CU(cudaMemcpyAsync(&h, bw, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        cudaFree(bw);
        ctx->bandwidthTiles = (h + TILE - 1) / TILE;
    }
    ctx->nTiles = gridFor(nc, TILE);
    TRY(devAlloc(ctx, &ctx->tileDesc, size_t(ctx->nTiles)));
    LAUNCH(k_tile_desc, gridFor(ctx->nTiles, 256), 256, ctx->ownerStart, ctx->extraStart, nc, ctx->nTiles,
           ctx->tileDesc);
    {
        int* cnt = nullptr;
        TRY(devAlloc(ctx, &cnt, 1));
        CU(cudaMemsetAsync(cnt, 0, sizeof(int), ctx->stream));
        LAUNCH(k_count_oversize, gridFor(ctx->nTiles, 256), 256, ctx->tileDesc, ctx->nTiles, cnt);
        CU(cudaMemcpyAsync(&ctx->nTilesOversize, cnt, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        cudaFree(cnt);
    }
    {
        std::vector<int> iota(ctx->nTiles);
        for (int t = 0; t < ctx->nTiles; ++t) iota[t] = t;
        TRY(devAlloc(ctx, &ctx->tileIota, iota.size()));
        CU(cudaMemcpyAsync(ctx->tileIota, iota.data(), iota.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    ctx->nTilesInterior = ctx->nTiles;
    ctx->nTilesCoupled = 0;
    if (ctx->nProcPatches > 0) {
        // tiles that touch a processor patch wait for the halos; all others can start at once
        int* flag = nullptr;
        TRY(devAlloc(ctx, &flag, size_t(ctx->nTiles)));
        LAUNCH(k_tile_coupled, ctx->nTiles, 128, ctx->extraStart, ctx->extra, nc, ctx->nTiles, flag);
        std::vector<int> h(ctx->nTiles), li, lc;
        CU(cudaMemcpyAsync(h.data(), flag, size_t(ctx->nTiles) * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        cudaFree(flag);
        for (int t = 0; t < ctx->nTiles; ++t) (h[t] ? lc : li).push_back(t);
        ctx->nTilesInterior = int(li.size());
        ctx->nTilesCoupled = int(lc.size());
        // a contiguous interior range (slab partitions in the tile-compact order) needs no list indirection, which
        // costs the cell kernel 4-5 % (profiles/r1_scaling.md)
        ctx->interiorRange0 = (!li.empty() && li.back() - li.front() + 1 == int(li.size())) ? li.front() : -1;
        TRY(devAlloc(ctx, &ctx->tileListInterior, li.size()));
        TRY(devAlloc(ctx, &ctx->tileListCoupled, lc.size()));
        if (!li.empty())
            CU(cudaMemcpyAsync(ctx->tileListInterior, li.data(), li.size() * sizeof(int), cudaMemcpyHostToDevice,
                               ctx->stream));
        if (!lc.empty())
            CU(cudaMemcpyAsync(ctx->tileListCoupled, lc.data(), lc.size() * sizeof(int), cudaMemcpyHostToDevice,
                               ctx->stream));
        TRY(devAlloc(ctx, &ctx->descInterior, li.size()));
        TRY(devAlloc(ctx, &ctx->descCoupled, lc.size()));
        if (!li.empty())
            LAUNCH(k_tile_desc_list, gridFor((long long)li.size(), 256), 256, ctx->tileDesc, ctx->tileListInterior,
                   int(li.size()), ctx->descInterior);
        if (!lc.empty())
            LAUNCH(k_tile_desc_list, gridFor((long long)lc.size(), 256), 256, ctx->tileDesc, ctx->tileListCoupled,
                   int(lc.size()), ctx->descCoupled);
    }
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(cursor);
    return SSAM_OK;
}

ssam_status uploadAoS(ssam_ctx* ctx, const double* host, int n, int nc, double* const* planes, int offset) {
    if (n == 0) return SSAM_OK;
    const int* perm = (offset == 0 && n == ctx->nCells) ? ctx->perm : nullptr;  // cell part in the library's order
    if (nc == 1 && !perm) {
        CU(cudaMemcpyAsync(planes[0] + offset, host, size_t(n) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        return SSAM_OK;
    }
    TRY(ensureStage(ctx, size_t(n) * nc));
    CU(cudaMemcpyAsync(ctx->stage, host, size_t(n) * nc * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    PlanePtrs pl;
    for (int k = 0; k < 16; ++k) pl.p[k] = k < nc ? planes[k] : nullptr;
    LAUNCH(k_aos_to_soa_n, gridFor((long long)n * nc, 256), 256, ctx->stage, n, nc, pl, offset, perm);
    return SSAM_OK;
}

ssam_status downloadAoS(ssam_ctx* ctx, double* host, int n, int nc, double* const* planes, int offset) {
    if (n == 0) return SSAM_OK;
    const int* perm = (offset == 0 && n == ctx->nCells) ? ctx->perm : nullptr;
    if (nc == 1 && !perm) {
        CU(cudaMemcpyAsync(host, planes[0] + offset, size_t(n) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        return SSAM_OK;
    }
    TRY(ensureStage(ctx, size_t(n) * nc));
    PlanePtrs pl;
    for (int k = 0; k < 16; ++k) pl.p[k] = k < nc ? planes[k] : nullptr;
    LAUNCH(k_soa_to_aos_n, gridFor((long long)n * nc, 256), 256, ctx->stage, n, nc, pl, offset, perm);
    CU(cudaMemcpyAsync(host, ctx->stage, size_t(n) * nc * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    // the staging buffer is reused by the next transfer
    CU(cudaStreamSynchronize(ctx->stream));
    return SSAM_OK;
}

ssam_status need(ssam_ctx* ctx, int f, const char* what) {
    CHECK(ctx->haveMesh, SSAM_ERR_INVALID, "mesh not set");
    if (!ctx->isSet[f]) return fail(ctx, SSAM_ERR_MISSING_FIELD, std::string("field not available: ") + what);
    return SSAM_OK;
}

ssam_status unitsShift(ssam_ctx* ctx, const ssam_conductivity& k, double* shift) {
    *shift = 0.0;
    if (k.kModel != SSAM_K_CUBIC) return SSAM_OK;
    *shift = -273.0;  // celsConv_, incompressibleTwoPhaseThermalMixture.C:226
    if (k.units == SSAM_UNITS_KELVIN)
        *shift = 0.0;
    else if (k.units != SSAM_UNITS_CELSIUS)
        return fail(ctx, SSAM_ERR_UNITS, "Unknown conductivity function units. Choose Kelvin or Celsius.");
    return SSAM_OK;
}

// dimensionedScalar sub-expressions the reference evaluates on the host, with the host's libm
ViscDerived deriveVisc(const ssam_viscosity_params& p) {
    ViscDerived d;
    std::memset(&d, 0, sizeof(d));
    d.invN = 1 / p.n;
    d.threeRho = 3.0 * p.rho;
    d.logA = std::log(p.A);
    d.log01 = std::log(0.1);
    d.Hcoef = p.a1 + (p.a2 * 1.571);
    d.TmTa = p.Tm - p.Ta;
    d.negC1 = -p.c1;
    d.invC2 = 1 / p.c2;
    d.b3LogY = p.b3 * std::log(0.1 / p.e0 + 1e-300);
    d.mu0OverRho = p.mu0 / p.rho;
    d.mMinus1 = p.m - 1;
    d.sqrt3 = std::sqrt(3.0);
    d.invE0 = 1.0 / p.e0;
    d.b2IsOne = (p.b2 == 1.0) ? 1 : 0;
    d.invC2IsOne = (d.invC2 == 1.0) ? 1 : 0;
    return d;
}

ssam_status checkViscModel(ssam_ctx* ctx, const ssam_viscosity_params* p) {
    CHECK(p != nullptr, SSAM_ERR_INVALID, "null viscosity params");
    CHECK(p->model >= SSAM_VISC_NEWTONIAN && p->model <= SSAM_VISC_NORTON_HOFF, SSAM_ERR_UNKNOWN_MODEL,
          "Unknown viscosityModel type");
    return SSAM_OK;
}

void fillMeshArgs(ssam_ctx* c, UpdateArgs& a) {
    std::memset(&a, 0, sizeof(a));
    a.nCells = c->nCells;
    a.nIF = c->nIF;
    a.nBnd = c->nBnd;
    a.geo = c->geo;
    a.owner = c->owner;
    a.neighbour = c->neighbour;
    a.ownerStart = c->ownerStart;
    a.extraStart = c->extraStart;
    a.extra = c->extra;
    a.tileDesc = c->tileDesc;
    a.V = c->V;
    a.magSf_b = c->magSf_b;
    a.deltaCoeffs_b = c->deltaCoeffs_b;
    a.bndFlags = c->bndFlags;
    a.Ux = c->plane[SSAM_F_U][0];
    a.Uy = c->plane[SSAM_F_U][1];
    a.Uz = c->plane[SSAM_F_U][2];
    for (int k = 0; k < 9; ++k) a.gradB[k] = c->gradB[k];
    a.epsFactor = std::sqrt(2.0 / 3.0);  // Foam::sqrt(2.0/3.0), fluid/TEqn.H:3
    a.srFactor = std::sqrt(2.0);         // viscosityModel::strainRate()
}

// ---- halo exchange --------------------------------------------------------------------------------------
// processorFvPatchField::evaluate for several fields at once: ONE pack kernel gathers field[faceCells] of
// every processor patch for all component planes, ONE NCCL group sends each (patch, plane) run and
// receives straight into the boundary part of the planes.
// stream memory operation: wait until *addr >= value without occupying an SM (driver entry point, no libcuda link)
typedef CUresult (*StreamWaitValue64Fn)(CUstream, CUdeviceptr, cuuint64_t, unsigned int);
StreamWaitValue64Fn streamWaitValue64() {
    static StreamWaitValue64Fn fn = []() -> StreamWaitValue64Fn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue64", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<StreamWaitValue64Fn>(p);
    }();
    return fn;
}

// One-time (per mesh) collective set-up of the peer-memory halo path: every rank allocates a receive buffer, the
// neighbours swap IPC handles and slot offsets through NCCL, and all ranks agree (all-reduce) on whether it works.
ssam_status p2pSetup(ssam_ctx* ctx) {
    auto& P = ctx->p2p;
    P.state = -1;
    // Round 1 ran both exchanges on a second stream beside the interior tiles; there NCCL was 12 us per step faster
    // (profiles/r1_scaling.md).  The round-2 schedule (updateFusedAsync) keeps the whole step on one stream and only
    // needs push / wait-flag / unpack pieces, which NCCL's send/recv pairs cannot be split into: peer memory is the
    // default, SSAM_HALO=nccl selects the NCCL exchange (also the fallback when a neighbour cannot be mapped).
    const char* env = std::getenv("SSAM_HALO");
    int ok = (env && std::string(env) == "nccl") ? 0 : 1;
    if (!streamWaitValue64()) ok = 0;
    NcclApi& N = nccl();
    const int nP = ctx->nProcPatches;
    struct Msg {
        cudaIpcMemHandle_t handle;
        int off, size;
        long long slotDoubles;  // size of one receive slot of the sender of this message
        int pad[12];
    };
    static_assert(sizeof(Msg) == 128, "handshake message size");
    std::vector<Msg> mine(static_cast<size_t>(nP)), theirs(static_cast<size_t>(nP));
    P.slotDoubles = size_t(ctx->nProcFaces) * 16;
    const size_t total = 32 + 2 * P.slotDoubles;  // 32 doubles = 256 B of flag words
    if (cudaMalloc(reinterpret_cast<void**>(&P.recvBuf), total * sizeof(double)) != cudaSuccess) ok = 0;
    cudaIpcMemHandle_t h;
    std::memset(&h, 0, sizeof(h));
    if (ok && cudaIpcGetMemHandle(&h, P.recvBuf) != cudaSuccess) ok = 0;
    cudaGetLastError();
    if (P.recvBuf) CU(cudaMemsetAsync(P.recvBuf, 0, total * sizeof(double), ctx->stream));
    TRY(devAlloc(ctx, &P.done, 1));
    CU(cudaMemsetAsync(P.done, 0, sizeof(unsigned), ctx->stream));
    std::vector<int> b0(static_cast<size_t>(nP) + 1, 0);
    for (int p = 0; p < nP; ++p) {
        const int patch = ctx->haloPatch[size_t(p)];
        b0[size_t(p)] = ctx->patchStart[size_t(patch)] - ctx->nIF;
        std::memset(&mine[size_t(p)], 0, sizeof(Msg));
        mine[size_t(p)].handle = h;
        mine[size_t(p)].off = ctx->haloOffsets[size_t(p)];
        mine[size_t(p)].size = ctx->haloOffsets[size_t(p) + 1] - ctx->haloOffsets[size_t(p)];
        mine[size_t(p)].slotDoubles = (long long)P.slotDoubles;
    }
    TRY(devAlloc(ctx, &P.patchB0, size_t(nP) + 1));
    CU(cudaMemcpyAsync(P.patchB0, b0.data(), (size_t(nP) + 1) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    // handshake with every neighbour (device staging buffers, one grouped send/recv)
    char* dmsg = nullptr;
    TRY(devAlloc(ctx, &dmsg, size_t(2 * nP) * sizeof(Msg)));
    CU(cudaMemcpyAsync(dmsg, mine.data(), size_t(nP) * sizeof(Msg), cudaMemcpyHostToDevice, ctx->stream));
    NC(N.GroupStart());
    for (int p = 0; p < nP; ++p) {
        const int peer = ctx->patchNbrRank[size_t(ctx->haloPatch[size_t(p)])];
        NC(N.Send(dmsg + size_t(p) * sizeof(Msg), sizeof(Msg), ncclChar, peer, ctx->comm, ctx->stream));
        NC(N.Recv(dmsg + size_t(nP + p) * sizeof(Msg), sizeof(Msg), ncclChar, peer, ctx->comm, ctx->stream));
    }
    NC(N.GroupEnd());
    CU(cudaMemcpyAsync(theirs.data(), dmsg + size_t(nP) * sizeof(Msg), size_t(nP) * sizeof(Msg), cudaMemcpyDeviceToHost,
                       ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int p = 0; p < nP && ok; ++p) {
        if (theirs[size_t(p)].size != mine[size_t(p)].size) {
            ok = 0;
            break;
        }
        // two patches towards the same neighbour share one mapping
        void* base = nullptr;
        for (int q = 0; q < p; ++q)
            if (ctx->patchNbrRank[size_t(ctx->haloPatch[size_t(q)])] == ctx->patchNbrRank[size_t(ctx->haloPatch[size_t(p)])])
                base = P.peerBase[q];
        if (!base) {
            if (cudaIpcOpenMemHandle(&base, theirs[size_t(p)].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                ok = 0;
                break;
            }
            P.peerMap[p] = base;
        }
        P.peerBase[p] = static_cast<double*>(base);
        P.peerOff[p] = theirs[size_t(p)].off;
        P.peerSlotDoubles[p] = size_t(theirs[size_t(p)].slotDoubles);
    }
    if (ctx->nRanks > 32) ok = 0;  // one flag word per neighbour rank, 32 words
    // every rank must take the same path
    int* dflag = reinterpret_cast<int*>(dmsg);
    CU(cudaMemcpyAsync(dflag, &ok, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    NC(N.AllReduce(dflag, dflag, 1, ncclInt, ncclMin, ctx->comm, ctx->stream));
    CU(cudaMemcpyAsync(&ok, dflag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(dmsg);
    if (!ok) {
        p2pFree(ctx);
        P.state = -1;
        return SSAM_OK;
    }
    P.state = 1;
    P.seq = 0;
    return SSAM_OK;
}

// push: gather field[faceCells] of every processor patch and store each run straight into the neighbour's receive slot
// over NVLink; the last CTA publishes the exchange number into the neighbours' flag words.  Returns that number.
ssam_status haloPushPeer(ssam_ctx* ctx, const PlanePtrs& pl, int nc, cudaStream_t stream, unsigned long long* seqOut) {
    auto& P = ctx->p2p;
    const unsigned long long seq = ++P.seq;
    const size_t slot = size_t(seq & 1ull);
    PeerLayout L;
    L.nPatches = ctx->nProcPatches;
    for (int p = 0; p <= ctx->nProcPatches; ++p) L.offsets[p] = ctx->haloOffsets[size_t(p)];
    for (int p = 0; p < ctx->nProcPatches; ++p) {
        double* base = P.peerBase[p];
        L.dst[p] = base + 32 + slot * P.peerSlotDoubles[p];  // the RECEIVER's slot size, not this rank's
        L.dstOff[p] = P.peerOff[p];
        // the neighbour's flag word for this rank: indexed by this rank's number modulo 32
        L.flag[p] = reinterpret_cast<unsigned long long*>(base) + (ctx->rank & 31);
    }
    const int grid = gridFor((long long)ctx->nProcFaces * nc, 256);
    k_halo_pack_peer<<<grid, 256, 0, stream>>>(ctx->haloSendCells, L, nc, pl, P.done, seq);
    ctx->launches++;
    CU(cudaGetLastError());
    ctx->haloBytesLastUpdate += size_t(ctx->nProcFaces) * size_t(nc) * sizeof(double);
    *seqOut = seq;
    return SSAM_OK;
}
// wait (stream memory operation, no spinning CTA) until every neighbour has published exchange `seq`, then move the
// received runs into the boundary part of the planes
ssam_status haloWaitUnpackPeer(ssam_ctx* ctx, const PlanePtrs& pl, int nc, unsigned long long seq, cudaStream_t stream) {
    auto& P = ctx->p2p;
    const size_t slot = size_t(seq & 1ull);
    StreamWaitValue64Fn wait = streamWaitValue64();
    for (int p = 0; p < ctx->nProcPatches; ++p) {
        const int peer = ctx->patchNbrRank[size_t(ctx->haloPatch[size_t(p)])];
        bool seen = false;
        for (int q = 0; q < p; ++q) seen |= ctx->patchNbrRank[size_t(ctx->haloPatch[size_t(q)])] == peer;
        if (seen) continue;
        const CUdeviceptr addr = reinterpret_cast<CUdeviceptr>(reinterpret_cast<unsigned long long*>(P.recvBuf) + (peer & 31));
        CHECK(wait(reinterpret_cast<CUstream>(stream), addr, seq, CU_STREAM_WAIT_VALUE_GEQ) == CUDA_SUCCESS, SSAM_ERR_CUDA,
              "cuStreamWaitValue64 failed");
    }
    HaloLayout H;
    H.nPatches = ctx->nProcPatches;
    for (int p = 0; p <= ctx->nProcPatches; ++p) H.offsets[p] = ctx->haloOffsets[size_t(p)];
    const int grid = gridFor((long long)ctx->nProcFaces * nc, 256);
    k_halo_unpack<<<grid, 256, 0, stream>>>(P.recvBuf + 32 + slot * P.slotDoubles, H, nc, pl, ctx->nCells, P.patchB0);
    ctx->launches++;
    CU(cudaGetLastError());
    return SSAM_OK;
}
ssam_status haloExchangePeer(ssam_ctx* ctx, const PlanePtrs& pl, int nc, cudaStream_t stream) {
    unsigned long long seq = 0;
    TRY(haloPushPeer(ctx, pl, nc, stream, &seq));
    return haloWaitUnpackPeer(ctx, pl, nc, seq, stream);
}

// collective, once per mesh: choose the halo transport (peer memory when every rank can map its neighbours)
// ---- in-process rank group -------------------------------------------------------------------------------------------
// own receive buffer, completion counter and patch table (what p2pSetup allocates before its handshake)
ssam_status p2pAllocOwn(ssam_ctx* ctx) {
    auto& P = ctx->p2p;
    if (P.recvBuf) return SSAM_OK;
    CU(cudaSetDevice(ctx->device));
    const int nP = ctx->nProcPatches;
    P.slotDoubles = size_t(ctx->nProcFaces) * 16;
    const size_t total = 32 + 2 * P.slotDoubles;
    TRY(devAlloc(ctx, &P.recvBuf, total));
    CU(cudaMemset(P.recvBuf, 0, total * sizeof(double)));
    TRY(devAlloc(ctx, &P.done, 1));
    CU(cudaMemset(P.done, 0, sizeof(unsigned)));
    std::vector<int> b0(size_t(nP) + 1, 0);
    for (int p = 0; p < nP; ++p) b0[size_t(p)] = ctx->patchStart[size_t(ctx->haloPatch[size_t(p)])] - ctx->nIF;
    TRY(devAlloc(ctx, &P.patchB0, size_t(nP) + 1));
    CU(cudaMemcpy(P.patchB0, b0.data(), (size_t(nP) + 1) * sizeof(int), cudaMemcpyHostToDevice));
    return SSAM_OK;
}
ssam_status p2pSetupLocal(ssam_ctx* ctx) {
    auto& P = ctx->p2p;
    CHECK(streamWaitValue64() != nullptr, SSAM_ERR_CUDA, "cuStreamWaitValue64 is not available");
    CHECK(ctx->nRanks <= 32, SSAM_ERR_INVALID, "at most 32 ranks in an in-process group");
    TRY(p2pAllocOwn(ctx));
    for (int p = 0; p < ctx->nProcPatches; ++p) {
        const int q = ctx->patchNbrRank[size_t(ctx->haloPatch[size_t(p)])];
        CHECK(q >= 0 && q < int(ctx->localPeers.size()) && ctx->localPeers[size_t(q)] != ctx, SSAM_ERR_COMM,
              "processor patch refers to a rank outside the in-process group");
        ssam_ctx* nb = ctx->localPeers[size_t(q)];
        CHECK(nb->haveMesh, SSAM_ERR_COMM, "in-process group: the neighbour rank has no mesh yet");
        int pq = -1;
        for (int k = 0; k < nb->nProcPatches; ++k)
            if (nb->patchNbrRank[size_t(nb->haloPatch[size_t(k)])] == ctx->rank) pq = k;
        CHECK(pq >= 0 && nb->haloOffsets[size_t(pq) + 1] - nb->haloOffsets[size_t(pq)] ==
                             ctx->haloOffsets[size_t(p) + 1] - ctx->haloOffsets[size_t(p)],
              SSAM_ERR_COMM, "processor patches of the two ranks do not match");
        {
            ssam_ctx* c2 = nb;  // allocate the neighbour's buffer if it has not reached its own set-up yet
            ssam_status st = p2pAllocOwn(c2);
            if (st != SSAM_OK) return fail(ctx, st, c2->err);
        }
        if (nb->device != ctx->device) {
            CU(cudaSetDevice(ctx->device));
            const cudaError_t e = cudaDeviceEnablePeerAccess(nb->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return fail(ctx, SSAM_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
            cudaGetLastError();
        }
        P.peerBase[p] = nb->p2p.recvBuf;
        P.peerOff[p] = nb->haloOffsets[size_t(pq)];
        P.peerSlotDoubles[p] = nb->p2p.slotDoubles;
    }
    CU(cudaSetDevice(ctx->device));
    P.state = 1;
    return SSAM_OK;
}
ssam_status ensureHaloTransport(ssam_ctx* ctx) {
    if (ctx->nProcPatches == 0 || ctx->p2p.state != 0) return SSAM_OK;
    if (!ctx->localPeers.empty()) return p2pSetupLocal(ctx);
    CHECK(ctx->comm != nullptr, SSAM_ERR_COMM, "mesh has processor patches but ssam_comm_init was not called");
    // every rank reaches its first exchange in the same call
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->commStream) CU(cudaStreamSynchronize(ctx->commStream));
    return p2pSetup(ctx);
}
ssam_status planesOf(ssam_ctx* ctx, const int* fields, int nFields, PlanePtrs& pl, int& nc) {
    nc = 0;
    for (int i = 0; i < nFields; ++i) {
        TRY(ensureField(ctx, fields[i]));
        for (int k = 0; k < ncompOf(fields[i]); ++k) {
            CHECK(nc < 16, SSAM_ERR_INVALID, "too many component planes in one halo exchange");
            pl.p[nc++] = ctx->plane[fields[i]][k];
        }
    }
    for (int k = nc; k < 16; ++k) pl.p[k] = nullptr;
    return SSAM_OK;
}

ssam_status haloExchangeFields(ssam_ctx* ctx, const int* fields, int nFields, cudaStream_t stream) {
    if (ctx->nProcPatches == 0 || nFields == 0) return SSAM_OK;
    CHECK(ctx->comm != nullptr || !ctx->localPeers.empty(), SSAM_ERR_COMM,
          "mesh has processor patches but ssam_comm_init was not called");
    CHECK(ctx->localPeers.empty() || !ctx->blocking, SSAM_ERR_INVALID,
          "in-process rank group: collective calls need ssam_set_blocking(ctx, 0) on every rank (enqueue all ranks, "
          "then synchronise)");
    TRY(ensureHaloTransport(ctx));
    PlanePtrs pl;
    int nc = 0;
    TRY(planesOf(ctx, fields, nFields, pl, nc));
    if (ctx->p2p.state == 1) return haloExchangePeer(ctx, pl, nc, stream);
    const size_t needBuf = size_t(ctx->nProcFaces) * 16;
    if (ctx->sendCap < needBuf) {
        CU(cudaStreamSynchronize(ctx->stream));
        if (ctx->commStream) CU(cudaStreamSynchronize(ctx->commStream));
        devFree(ctx->sendBuf);
        TRY(devAlloc(ctx, &ctx->sendBuf, needBuf));
        ctx->sendCap = needBuf;
    }
    HaloLayout L;
    L.nPatches = ctx->nProcPatches;
    for (int p = 0; p <= ctx->nProcPatches; ++p) L.offsets[p] = ctx->haloOffsets[p];
    k_halo_pack<<<gridFor((long long)ctx->nProcFaces * nc, 256), 256, 0, stream>>>(ctx->haloSendCells, L, nc, pl,
                                                                                    ctx->sendBuf);
    ctx->launches++;
    CU(cudaGetLastError());
    NcclApi& N = nccl();
    NC(N.GroupStart());
    for (int p = 0; p < ctx->nProcPatches; ++p) {
        const int patch = ctx->haloPatch[p];
        const int peer = ctx->patchNbrRank[patch];
        const int off = ctx->haloOffsets[p], np = ctx->haloOffsets[p + 1] - off;
        const int b0 = ctx->patchStart[patch] - ctx->nIF;
        for (int k = 0; k < nc; ++k) {
            NC(N.Send(ctx->sendBuf + size_t(nc) * off + size_t(k) * np, np, ncclFloat64, peer, ctx->comm, stream));
            NC(N.Recv(pl.p[k] + ctx->nCells + b0, np, ncclFloat64, peer, ctx->comm, stream));
        }
    }
    NC(N.GroupEnd());
    ctx->launches++;  // one fused NCCL p2p kernel per group
    ctx->haloBytesLastUpdate += size_t(ctx->nProcFaces) * size_t(nc) * sizeof(double);
    return SSAM_OK;
}
ssam_status haloExchangeField(ssam_ctx* ctx, int field) { return haloExchangeFields(ctx, &field, 1, ctx->stream); }

// The same exchange for data that is not a field: gather plSend[k][faceCells], deliver into plRecv[k][nCells + b]
// (pass plRecv[k] = array - nCells to receive into a boundary-sized array).  Used once per mesh for the neighbour
// cell centres (processorPolyPatch::neighbFaceCellCentres()).
ssam_status haloExchangePlanes(ssam_ctx* ctx, const PlanePtrs& plSend, const PlanePtrs& plRecv, int nc, cudaStream_t stream) {
    if (ctx->nProcPatches == 0) return SSAM_OK;
    CHECK(ctx->comm != nullptr || !ctx->localPeers.empty(), SSAM_ERR_COMM,
          "mesh has processor patches but ssam_comm_init was not called");
    CHECK(ctx->localPeers.empty() || !ctx->blocking, SSAM_ERR_INVALID,
          "in-process rank group: collective calls need ssam_set_blocking(ctx, 0) on every rank (enqueue all ranks, "
          "then synchronise)");
    TRY(ensureHaloTransport(ctx));
    if (ctx->p2p.state == 1) {
        unsigned long long seq = 0;
        TRY(haloPushPeer(ctx, plSend, nc, stream, &seq));
        return haloWaitUnpackPeer(ctx, plRecv, nc, seq, stream);
    }
    const size_t needBuf = size_t(ctx->nProcFaces) * 16;
    if (ctx->sendCap < needBuf) {
        CU(cudaStreamSynchronize(ctx->stream));
        devFree(ctx->sendBuf);
        TRY(devAlloc(ctx, &ctx->sendBuf, needBuf));
        ctx->sendCap = needBuf;
    }
    HaloLayout L;
    L.nPatches = ctx->nProcPatches;
    for (int p = 0; p <= ctx->nProcPatches; ++p) L.offsets[p] = ctx->haloOffsets[p];
    k_halo_pack<<<gridFor((long long)ctx->nProcFaces * nc, 256), 256, 0, stream>>>(ctx->haloSendCells, L, nc, plSend,
                                                                                    ctx->sendBuf);
    ctx->launches++;
    CU(cudaGetLastError());
    NcclApi& N = nccl();
    NC(N.GroupStart());
    for (int p = 0; p < ctx->nProcPatches; ++p) {
        const int patch = ctx->haloPatch[p];
        const int peer = ctx->patchNbrRank[patch];
        const int off = ctx->haloOffsets[p], np = ctx->haloOffsets[p + 1] - off;
        const int b0 = ctx->patchStart[patch] - ctx->nIF;
        for (int k = 0; k < nc; ++k) {
            NC(N.Send(ctx->sendBuf + size_t(nc) * off + size_t(k) * np, np, ncclFloat64, peer, ctx->comm, stream));
            NC(N.Recv(plRecv.p[k] + ctx->nCells + b0, np, ncclFloat64, peer, ctx->comm, stream));
        }
    }
    NC(N.GroupEnd());
    ctx->launches++;
    return SSAM_OK;
}

ssam_status ensureFricSums(ssam_ctx* ctx, int patch) {
    if (ctx->fricPatchCached == patch) return SSAM_OK;
    if (!ctx->fricSums) TRY(devAlloc(ctx, &ctx->fricSums, 8));
    const int b0 = ctx->patchStart[patch] - ctx->nIF;
    LAUNCH(k_patch_sums, 1, 32, b0, ctx->patchSize[patch], ctx->Cpl[0], ctx->Cpl[1], ctx->Cpl[2], ctx->magSf_b,
           ctx->nCells, ctx->fricSums);
    if (!ctx->localPeers.empty() && ctx->nRanks > 1) {
        // in-process group: the ranks' sums (constantSlipStick.C:118-131, serial face order, area starting at VSMALL on
        // every rank like the reference) formed on the host from the caller's geometry and added in rank order
        double tot[4] = {0, 0, 0, 0};
        for (size_t r = 0; r < ctx->localPeers.size(); ++r) {
            const ssam_ctx* q = ctx->localPeers[r];
            CHECK(q->haveMesh && patch < q->nPatches, SSAM_ERR_COMM, "in-process group: neighbour rank has no such patch");
            double c[3] = {0.0, 0.0, 0.0}, a = 1e-300;
            const int b0 = q->patchStart[size_t(patch)] - q->nIF;
            for (int i = 0; i < q->patchSize[size_t(patch)]; ++i) {
                const double ms = q->hostMagSfB[size_t(b0 + i)];
                for (int k = 0; k < 3; ++k) c[k] += q->hostCfB[3 * size_t(b0 + i) + size_t(k)] * ms;
                a += ms;
            }
            const double loc[4] = {c[0], c[1], c[2], a};
            for (int k = 0; k < 4; ++k) tot[k] = r == 0 ? loc[k] : tot[k] + loc[k];
        }
        for (int k = 0; k < 4; ++k) ctx->fricHost[k] = tot[k];
        // by kernel argument: a pageable cudaMemcpyAsync would synchronise the stream, which may be parked on a wait
        // for a rank whose work this host thread has not enqueued yet
        LAUNCH(k_set4, 1, 1, ctx->fricSums + 4, tot[0], tot[1], tot[2], tot[3]);
    } else if (ctx->comm && ctx->nRanks > 1) {
        // reduce(patchCenter_, sumOp<vector>()); reduce(patchArea_, sumOp<scalar>())  (constantSlipStick.C:134-135)
        // NB every rank starts its area at VSMALL like the reference does.
        NC(nccl().AllReduce(ctx->fricSums, ctx->fricSums + 4, 4, ncclFloat64, ncclSum, ctx->comm, ctx->stream));
        ctx->launches++;
    } else {
        CU(cudaMemcpyAsync(ctx->fricSums + 4, ctx->fricSums, 4 * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    ctx->fricPatchCached = patch;
    return SSAM_OK;
}

ssam_status launchQfric(ssam_ctx* ctx, const ssam_stick_params& p) {
    CHECK(p.fricPatch < ctx->nPatches, SSAM_ERR_INVALID, "fricPatch out of range (findPatchID returned -1?)");
    CHECK(p.model == SSAM_STICK_CONSTANT || p.model == SSAM_STICK_EXPONENTIAL, SSAM_ERR_UNKNOWN_MODEL,
          "Unknown stickModel type");
    TRY(need(ctx, SSAM_F_SIGMAY, "sigmay (NortonHoff registers none, SURVEY D-1)"));
    TRY(need(ctx, SSAM_F_ALPHA1, "alpha1"));
    TRY(need(ctx, SSAM_F_P, "p"));
    TRY(ensureField(ctx, SSAM_F_QFRIC));
    if (ctx->qfricZeroedFor != p.fricPatch) {
        // fricCell_ = alpha_*0.0 everywhere except the patch cells (:79-92): zero-filled once per patch
        CU(cudaMemsetAsync(ctx->plane[SSAM_F_QFRIC][0], 0, size_t(ctx->nTot) * sizeof(double), ctx->stream));
        ctx->qfricZeroedFor = p.fricPatch;
    }
    TRY(ensureFricSums(ctx, p.fricPatch));
    const int n = ctx->patchSize[p.fricPatch];
    LAUNCH(k_qfric, gridFor(n, 128), 128, p, ctx->patchStart[p.fricPatch], n, ctx->owner, ctx->Cpl[0], ctx->Cpl[1],
           ctx->Cpl[2], ctx->fricSums + 4, ctx->plane[SSAM_F_ALPHA1][0], ctx->plane[SSAM_F_SIGMAY][0],
           ctx->plane[SSAM_F_P][0], ctx->plane[SSAM_F_QFRIC][0]);
    ctx->isSet[SSAM_F_QFRIC] = true;
    return SSAM_OK;
}

template <int MODE>
ssam_status launchCellsBoundary(ssam_ctx* ctx, const UpdateArgs& a, int model1, bool boundaryToo) {
    const int gc = gridFor(ctx->nCells, 256), gb = gridFor(ctx->nBnd, 128);
#define SSAM_DISPATCH(M)                                                    \
    case M:                                                                 \
        if (!boundaryToo) {                                                 \
            LAUNCH((k_cells<M, MODE>), gc, 256, a);                         \
        } else {                                                            \
            LAUNCH((k_boundary<M, MODE>), gb, 128, a);                      \
        }                                                                   \
        break;
    switch (model1) {
        SSAM_DISPATCH(SSAM_VISC_NEWTONIAN)
        SSAM_DISPATCH(SSAM_VISC_SHEPPARD_WRIGHT)
        SSAM_DISPATCH(SSAM_VISC_FKS)
        SSAM_DISPATCH(SSAM_VISC_NORTON_HOFF)
        default:
            return fail(ctx, SSAM_ERR_UNKNOWN_MODEL, "Unknown viscosityModel type");
    }
#undef SSAM_DISPATCH
    return SSAM_OK;
}

template <int M, bool LIST, bool FITS>
ssam_status launchTiledOne(ssam_ctx* ctx, const UpdateArgs& a0, const int* list, int count, cudaStream_t stream) {
    if (count == 0) return SSAM_OK;
    UpdateArgs a = a0;
    a.tileList = list;
    a.tileCount = count;
    if (LIST) {
        if (ctx->tileListInterior && list >= ctx->tileListInterior && list < ctx->tileListInterior + ctx->nTilesInterior) {
            a.tileDesc = ctx->descInterior + (list - ctx->tileListInterior);
        } else if (ctx->tileListCoupled && list >= ctx->tileListCoupled && list < ctx->tileListCoupled + ctx->nTilesCoupled) {
            a.tileDesc = ctx->descCoupled + (list - ctx->tileListCoupled);
        } else if (list == ctx->tileOrder) {
            a.tileDesc = ctx->descOrder;
        } else if (ctx->pipeList && list >= ctx->pipeList && list < ctx->pipeList + ctx->nTiles) {
            a.tileDesc = ctx->descPipe + (list - ctx->pipeList);  // a range of the caller-order list
        } else {
            a.tileDesc = ctx->tileDesc + (list - ctx->tileIota);  // a range of the identity list
        }
    }
    static bool configured[64] = {false};  // per device: the attribute belongs to the device's context
    const int smem = int(sizeof(TileSmem));
    if (!configured[ctx->device & 63]) {
        CU(cudaFuncSetAttribute((k_cells_tiled<M, LIST, FITS>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured[ctx->device & 63] = true;
    }
    k_cells_tiled<M, LIST, FITS><<<count, TILE, smem, stream>>>(a);
    ctx->launches++;
    CU(cudaGetLastError());
    return SSAM_OK;
}
// persistent, self-pipelined variant: one CTA per resident slot, looping over the work list
template <int M, bool FITS>
ssam_status launchPipeOne(ssam_ctx* ctx, const UpdateArgs& a0, const int* list, int count, cudaStream_t stream) {
    if (count == 0) return SSAM_OK;
    UpdateArgs a = a0;
    a.tileList = list;
    a.tileCount = count;
    if (list) {
        if (ctx->tileListInterior && list >= ctx->tileListInterior && list < ctx->tileListInterior + ctx->nTilesInterior) {
            a.tileDesc = ctx->descInterior + (list - ctx->tileListInterior);
        } else if (ctx->tileListCoupled && list >= ctx->tileListCoupled && list < ctx->tileListCoupled + ctx->nTilesCoupled) {
            a.tileDesc = ctx->descCoupled + (list - ctx->tileListCoupled);
        } else if (list == ctx->tileOrder) {
            a.tileDesc = ctx->descOrder;
        } else if (ctx->pipeList && list >= ctx->pipeList && list < ctx->pipeList + ctx->nTiles) {
            a.tileDesc = ctx->descPipe + (list - ctx->pipeList);  // a range of the caller-order list
        } else {
            a.tileDesc = ctx->tileDesc + (list - ctx->tileIota);  // a range of the identity list
        }
    }
    static int residentCtas[64] = {0};  // per device: SMs x CTAs per SM of this instantiation
    const int smem = int(sizeof(TileSmem));
    int& slots = residentCtas[ctx->device & 63];
    if (slots == 0) {
        CU(cudaFuncSetAttribute(k_cells_pipe<M, FITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int perSm = 0, sms = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_cells_pipe<M, FITS>, TILE, size_t(smem)));
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
        CHECK(perSm > 0 && sms > 0, SSAM_ERR_CUDA, "k_cells_pipe does not fit on this device");
        slots = perSm * sms;
    }
    int grid = std::min(count, slots);
    if (ctx->pipeGridCap > 0) grid = std::min(grid, ctx->pipeGridCap);
    // the persistent kernel issues a tile's copies a whole chain phase ahead of their use: the L2 prefetch that
    // k_cells_tiled needs only adds issue work here (measured, profiles/r2_k_cells_pipe.md) -- opt-in via SSAM_PF_PIPE
    static const int pfPipe = std::getenv("SSAM_PF_PIPE") ? std::atoi(std::getenv("SSAM_PF_PIPE")) : 0;
    if (pfPipe > 0) {
        a.pfDistU = pfPipe + grid + (a.pfDistU - a.pfDist);
        a.pfDist = pfPipe + grid;
    } else {
        a.pfDist = a.pfDistU = 0;
    }
    // SSAM_PIPE_DYNAMIC=0: static round-robin instead of the atomic work counter; SSAM_PIPE_STAGGER_NS: start stagger
    static const int dynamicSched = std::getenv("SSAM_PIPE_DYNAMIC") ? std::atoi(std::getenv("SSAM_PIPE_DYNAMIC")) : 1;
    static const int staggerNs = std::getenv("SSAM_PIPE_STAGGER_NS") ? std::atoi(std::getenv("SSAM_PIPE_STAGGER_NS")) : 6000;
    a.sched = nullptr;
    if (dynamicSched && grid < count) {
        if (!ctx->pipeSched) {
            TRY(devAlloc(ctx, &ctx->pipeSched, 16));
            CU(cudaMemsetAsync(ctx->pipeSched, 0, 16 * sizeof(int), stream));
        }
        a.sched = ctx->pipeSched + 2 * (ctx->pipeSchedNext++ & 7);
    }
    a.staggerNs = grid < count ? staggerNs : 0;
    k_cells_pipe<M, FITS><<<grid, TILE, smem, stream>>>(a);
    ctx->launches++;
    CU(cudaGetLastError());
    return SSAM_OK;
}
ssam_status launchCellsPipe(ssam_ctx* ctx, const UpdateArgs& a, int model1, const int* list, int count,
                            cudaStream_t stream) {
#define SSAM_PIPE(M)                                                                         \
    case M:                                                                                  \
        return ctx->nTilesOversize == 0 ? launchPipeOne<M, true>(ctx, a, list, count, stream) \
                                        : launchPipeOne<M, false>(ctx, a, list, count, stream);
    switch (model1) {
        SSAM_PIPE(SSAM_VISC_NEWTONIAN)
        SSAM_PIPE(SSAM_VISC_SHEPPARD_WRIGHT)
        SSAM_PIPE(SSAM_VISC_FKS)
        SSAM_PIPE(SSAM_VISC_NORTON_HOFF)
        SSAM_PIPE(TILED_GRAD)
        SSAM_PIPE(TILED_STRAIN)
    }
#undef SSAM_PIPE
    return fail(ctx, SSAM_ERR_UNKNOWN_MODEL, "Unknown viscosityModel type");
}

ssam_status launchCellsTiled(ssam_ctx* ctx, const UpdateArgs& a, int model1, const int* list, int count,
                             cudaStream_t stream) {
    if (ctx->cellKernel == 2) return launchCellsPipe(ctx, a, model1, list, count, stream);
    // SSAM_FITS=0 keeps the general gather (with global-memory fallbacks) even when every tile fits (experiments)
    static const bool allowFits = !(std::getenv("SSAM_FITS") && std::atoi(std::getenv("SSAM_FITS")) == 0);
    const bool fits = allowFits && ctx->nTilesOversize == 0;
#define SSAM_TILED(M)                                                                           \
    case M:                                                                                     \
        if (fits)                                                                               \
            return list ? launchTiledOne<M, true, true>(ctx, a, list, count, stream)            \
                        : launchTiledOne<M, false, true>(ctx, a, nullptr, count, stream);       \
        return list ? launchTiledOne<M, true, false>(ctx, a, list, count, stream)               \
                    : launchTiledOne<M, false, false>(ctx, a, nullptr, count, stream);
    switch (model1) {
        SSAM_TILED(SSAM_VISC_NEWTONIAN)
        SSAM_TILED(SSAM_VISC_SHEPPARD_WRIGHT)
        SSAM_TILED(SSAM_VISC_FKS)
        SSAM_TILED(SSAM_VISC_NORTON_HOFF)
        SSAM_TILED(TILED_GRAD)
        SSAM_TILED(TILED_STRAIN)
    }
#undef SSAM_TILED
    return fail(ctx, SSAM_ERR_UNKNOWN_MODEL, "Unknown viscosityModel type");
}

// fvc::grad(U) pieces: MODE_GRAD writes GRADU, MODE_STRAIN writes factor*mag(symm(grad U)) into outField
ssam_status gradOrStrain(ssam_ctx* ctx, bool strain, double factor, int outField) {
    TRY(need(ctx, SSAM_F_U, "U"));
    UpdateArgs a;
    fillMeshArgs(ctx, a);
    a.pfDist = ctx->pfDist;
    a.pfDistU = a.pfDist + ctx->pfBandwidthTiles;
    const bool tiled = ctx->cellKernel >= 1;
    if (strain) {
        TRY(ensureField(ctx, outField));
        a.strainOut = ctx->plane[outField][0];
        a.strainFactor = factor;
        if (tiled) {
            TRY(launchCellsTiled(ctx, a, TILED_STRAIN, ctx->tileOrder, ctx->nTiles, ctx->stream));
        } else {
            TRY((launchCellsBoundary<MODE_STRAIN>(ctx, a, SSAM_VISC_NEWTONIAN, false)));
        }
        if (ctx->nProcPatches) TRY(haloExchangeField(ctx, outField));
        TRY((launchCellsBoundary<MODE_STRAIN>(ctx, a, SSAM_VISC_NEWTONIAN, true)));
        ctx->isSet[outField] = true;
    } else {
        TRY(ensureField(ctx, SSAM_F_GRADU));
        for (int k = 0; k < 9; ++k) a.gradU[k] = ctx->plane[SSAM_F_GRADU][k];
        if (tiled) {
            TRY(launchCellsTiled(ctx, a, TILED_GRAD, ctx->tileOrder, ctx->nTiles, ctx->stream));
        } else {
            TRY((launchCellsBoundary<MODE_GRAD>(ctx, a, SSAM_VISC_NEWTONIAN, false)));
        }
        if (ctx->nProcPatches) TRY(haloExchangeField(ctx, SSAM_F_GRADU));
        TRY((launchCellsBoundary<MODE_GRAD>(ctx, a, SSAM_VISC_NEWTONIAN, true)));
        ctx->isSet[SSAM_F_GRADU] = true;
    }
    return SSAM_OK;
}

ssam_status viscosityCorrect(ssam_ctx* ctx, int phase, const ssam_viscosity_params* p) {
    CHECK(phase == 1 || phase == 2, SSAM_ERR_INVALID, "phase must be 1 or 2");
    TRY(checkViscModel(ctx, p));
    const int nuField = phase == 1 ? SSAM_F_NU1 : SSAM_F_NU2;
    TRY(ensureField(ctx, nuField));
    ViscStage s;
    s.p = *p;
    s.d = deriveVisc(*p);
    s.T = s.eps = s.sr = nullptr;
    s.sigmaf = s.sigmay = nullptr;
    s.nu = ctx->plane[nuField][0];
    if (p->model == SSAM_VISC_SHEPPARD_WRIGHT || p->model == SSAM_VISC_FKS) {
        TRY(need(ctx, SSAM_F_T, "temp"));
        TRY(need(ctx, SSAM_F_EPSDOT, "epsilonDotEq"));
        TRY(ensureField(ctx, SSAM_F_SIGMAF));
        TRY(ensureField(ctx, SSAM_F_SIGMAY));
        s.T = ctx->plane[SSAM_F_T][0];
        s.eps = ctx->plane[SSAM_F_EPSDOT][0];
        s.sigmaf = ctx->plane[SSAM_F_SIGMAF][0];
        s.sigmay = ctx->plane[SSAM_F_SIGMAY][0];
    } else if (p->model == SSAM_VISC_NORTON_HOFF) {
        // calcNu() calls strainRate() = sqrt(2)*mag(symm(fvc::grad(U_)))  (NortonHoff.C:62)
        TRY(gradOrStrain(ctx, true, std::sqrt(2.0), SSAM_F_STRAINRATE));
        s.sr = ctx->plane[SSAM_F_STRAINRATE][0];
    }
    LAUNCH((k_map<ViscStage>), std::min(gridFor(ctx->nTot, 256), 148 * 32), 256, ctx->nTot, s);
    ctx->isSet[nuField] = true;
    if (nuField == SSAM_F_NU2) ctx->nu2Filled = false;
    if (s.sigmaf) ctx->isSet[SSAM_F_SIGMAF] = ctx->isSet[SSAM_F_SIGMAY] = true;
    return SSAM_OK;
}

int mapGrid(const ssam_ctx* c) { return std::min(gridFor(c->nTot, 256), 148 * 32); }

// Traversal order of the tiles.  The cell kernel re-reads, from L2, the face records and U values that the
// tile `bandwidth` cells earlier brought in; in natural order that reuse distance is the cell bandwidth
// of the mesh (nx*ny cells on a blockMesh box: 70 MB of traffic for a 512x512 plane, more than L2 keeps).
// A breadth-first (Cuthill-McKee) order of the tile graph walks the mesh in wavefronts, which cuts the
// reuse distance to one wavefront.  Pure integer preprocessing on the host arrays; cells and faces are NOT
// renumbered, so the per-cell summation order - and with it every result bit - is unchanged.
ssam_status buildTileOrder(ssam_ctx* ctx, const ssam_mesh_desc* m) {
    devFree(ctx->tileOrder);
    const long long reuseBytes = (long long)ctx->pfBandwidthTiles * TILE * 270;  // reuse distance of the heavy gathers
    if (reuseBytes < (40ll << 20) || ctx->nTiles < 4) return SSAM_OK;  // natural order already fits L2
    static const int bfs = std::getenv("SSAM_BFS_ORDER") ? std::atoi(std::getenv("SSAM_BFS_ORDER")) : 1;
    if (!bfs) return SSAM_OK;
    const int nT = ctx->nTiles;
    std::vector<std::vector<int>> adj(nT);
    for (int f = 0; f < m->nInternalFaces; ++f) {
        const int ta = m->owner[f] / TILE, tb = m->neighbour[f] / TILE;
        if (ta == tb) continue;
        if (adj[ta].empty() || adj[ta].back() != tb) adj[ta].push_back(tb);
        if (adj[tb].empty() || adj[tb].back() != ta) adj[tb].push_back(ta);
    }
    for (auto& v : adj) {
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
    }
    std::vector<int> order;
    order.reserve(nT);
    std::vector<char> seen(nT, 0);
    for (int root = 0; root < nT; ++root) {
        if (seen[root]) continue;
        seen[root] = 1;
        order.push_back(root);
        for (size_t head = order.size() - 1; head < order.size(); ++head)
            for (int t : adj[order[head]])
                if (!seen[t]) {
                    seen[t] = 1;
                    order.push_back(t);
                }
    }
    TRY(devAlloc(ctx, &ctx->tileOrder, size_t(nT)));
    CU(cudaMemcpyAsync(ctx->tileOrder, order.data(), size_t(nT) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    TRY(devAlloc(ctx, &ctx->descOrder, size_t(nT)));
    LAUNCH(k_tile_desc_list, gridFor(nT, 256), 256, ctx->tileDesc, ctx->tileOrder, nT, ctx->descOrder);
    CU(cudaStreamSynchronize(ctx->stream));
    return SSAM_OK;
}

// ---- tile-compact cell layout ---------------------------------------------------------------------------------
// The tiled kernel finds a face neighbour in shared memory only when it lies in the same 256 consecutive cells.
// In blockMesh order such a tile is one x-row and 4 of a cell's 6 neighbours are elsewhere.  The library may
// therefore keep the cells in its own order: runs of RUN consecutive caller cells are clustered greedily into
// tiles of TILE/RUN runs that are mesh neighbours of each other (heaviest connection first; ties: the run
// adjacent to the oldest member, then the lowest label) -- on a blockMesh box 64x2x2 bricks, 2 remote neighbours
// per cell.  Runs stay contiguous, so uploads and downloads (which apply the permutation) remain coalesced.
// Faces keep the caller's owner/neighbour roles and every cell still visits its faces in the caller's ascending
// face order (k_extra_sort's key), so all results are bit-identical to the caller-order layout.
constexpr int RUN = 64, RUNS_PER_TILE = TILE / RUN;

struct PermutedMesh {
    std::vector<int> perm, inv, owner, neighbour, faceKey;
    std::vector<double> Sf, w, V, C;
};

// Order of the tiles of the tile-compact layout in memory (= the order the cell kernel walks them).  SSAM_TILE_ORDER:
//   -1 (default) automatic: seed order, unless the heavy out-of-tile gathers would then be further away than L2
//      holds (planes wider than ~256 x 512 cells), in which case the pencil order;
//    0 seed order (tiles in the order of their first run in the caller's numbering);
//    1 Morton order of the tile centroids within 16 coarse caller-order buckets (measured: never the best);
//    2 pencil order (long tile direction slowest).
// Measured on one B200 (profiles/r2_tile_order.md): 256x256 planes seed 0.78-0.80 / pencil 0.75-0.77 of HBM; 512-wide
// planes seed 0.68 / pencil 0.73.
int tileOrderModeFromEnv() {
    const char* e = std::getenv("SSAM_TILE_ORDER");
    return e ? std::atoi(e) : -1;
}

// -> perm (library position -> caller cell); returns false when clustering does not pay
bool computeCompactOrder(int layoutMode, int tileOrderMode, const ssam_mesh_desc* m, std::vector<int>& perm,
                         double& remotePerCellCaller, double& remotePerCellCompact) {
    const int nC = m->nCells, nIF = m->nInternalFaces;
    const int nB = (nC + RUN - 1) / RUN;
    std::vector<std::vector<int>> nb(nB);
    long long remoteCaller = 0;
    for (int f = 0; f < nIF; ++f) {
        const int o = m->owner[f], n = m->neighbour[f];
        if (o / TILE != n / TILE) remoteCaller += 2;
        const int bo = o / RUN, bn = n / RUN;
        if (bo == bn) continue;
        nb[bo].push_back(bn);
        nb[bn].push_back(bo);
    }
    std::vector<int> adjStart(nB + 1, 0), adjTo, adjW;
    for (int b = 0; b < nB; ++b) {
        std::vector<int>& v = nb[b];
        std::sort(v.begin(), v.end());
        for (size_t i = 0; i < v.size();) {
            size_t j = i;
            while (j < v.size() && v[j] == v[i]) ++j;
            adjTo.push_back(v[i]);
            adjW.push_back(int(j - i));
            i = j;
        }
        adjStart[b + 1] = int(adjTo.size());
        std::vector<int>().swap(v);
    }
    std::vector<char> assigned(nB, 0);
    std::vector<int> full, partial;  // runs of complete tiles first, leftovers last (keeps tiles aligned)
    struct Cand { int run, w, age; };
    std::vector<Cand> cand;
    for (int seed = 0; seed < nB; ++seed) {
        if (assigned[seed]) continue;
        int members[RUNS_PER_TILE];
        int nMem = 0;
        cand.clear();
        auto add = [&](int b) {
            assigned[b] = 1;
            members[nMem++] = b;
            for (int k = adjStart[b]; k < adjStart[b + 1]; ++k) {
                const int j = adjTo[k];
                if (assigned[j]) continue;
                bool found = false;
                for (Cand& q : cand)
                    if (q.run == j) {
                        q.w += adjW[k];
                        found = true;
                        break;
                    }
                if (!found) cand.push_back({j, adjW[k], nMem});
            }
        };
        add(seed);
        while (nMem < RUNS_PER_TILE && !cand.empty()) {
            size_t best = 0;
            for (size_t q = 1; q < cand.size(); ++q) {
                const Cand &x = cand[q], &y = cand[best];
                if (x.w > y.w || (x.w == y.w && (x.age < y.age || (x.age == y.age && x.run < y.run)))) best = q;
            }
            const int j = cand[best].run;
            cand.erase(cand.begin() + long(best));
            // the last (short) run of the mesh would misalign every later tile: it goes to the leftovers
            if ((j + 1) * RUN > nC) continue;
            add(j);
        }
        std::sort(members, members + nMem);
        const bool whole = nMem == RUNS_PER_TILE && (members[nMem - 1] + 1) * RUN <= nC;
        for (int k = 0; k < nMem; ++k) (whole ? full : partial).push_back(members[k]);
    }
    std::sort(partial.begin(), partial.end());
    // Order of the complete tiles in memory = order in which the persistent kernel walks them.  In seed order the
    // bricks of a blockMesh box follow x, y, z, so a tile's z-neighbours are one whole layer of bricks away
    // (512x512 planes: 140 MB of traffic, more than the 126 MB L2 keeps -> the out-of-tile gathers miss).  A Morton
    // (Z-order) curve over the tile centroids bounds the reuse distance by the surface of a sub-cube instead.
    // Integer/geometry preprocessing only: per-cell face order, and with it every result bit, is unchanged.
    if (tileOrderMode < 0) {
        // automatic: 97th percentile of the tile distance of the internal faces in seed order = how far back the heavy
        // gathers reach; beyond ~48 MB of per-tile traffic (64 KB per tile) they miss L2
        std::vector<int> tileOfRun(size_t(nB), 0);
        for (size_t i = 0; i < full.size(); ++i) tileOfRun[size_t(full[i])] = int(i / RUNS_PER_TILE);
        const int nFull = int(full.size() / RUNS_PER_TILE);
        for (size_t i = 0; i < partial.size(); ++i) tileOfRun[size_t(partial[i])] = nFull + int(i / RUNS_PER_TILE);
        std::vector<int> dist;
        dist.reserve(size_t(nIF) / 16 + 1);
        for (int f = 0; f < nIF; f += 16)
            dist.push_back(std::abs(tileOfRun[size_t(m->neighbour[f] / RUN)] - tileOfRun[size_t(m->owner[f] / RUN)]));
        tileOrderMode = 0;
        if (!dist.empty()) {
            const size_t k = std::min(dist.size() - 1, size_t(0.97 * double(dist.size())));
            std::nth_element(dist.begin(), dist.begin() + long(k), dist.end());
            if ((long long)dist[k] * 65536ll > (48ll << 20)) tileOrderMode = 2;
        }
    }
    if (tileOrderMode == 1 && full.size() >= size_t(8 * RUNS_PER_TILE)) {
        const size_t nT = full.size() / RUNS_PER_TILE;
        std::vector<double> cen(nT * 3, 0.0);
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (size_t t = 0; t < nT; ++t) {
            for (int r = 0; r < RUNS_PER_TILE; ++r) {
                const int b = full[t * RUNS_PER_TILE + size_t(r)];
                // centroid of a run ~ mean of its first and last cell centre (runs are mesh lines)
                const int c0 = b * RUN, c1 = std::min(nC, (b + 1) * RUN) - 1;
                for (int k = 0; k < 3; ++k)
                    cen[t * 3 + size_t(k)] += 0.5 * (m->C[size_t(c0) * 3 + size_t(k)] + m->C[size_t(c1) * 3 + size_t(k)]);
            }
            for (int k = 0; k < 3; ++k) {
                cen[t * 3 + size_t(k)] /= RUNS_PER_TILE;
                lo[k] = std::min(lo[k], cen[t * 3 + size_t(k)]);
                hi[k] = std::max(hi[k], cen[t * 3 + size_t(k)]);
            }
        }
        const double ext = std::max({hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2], 1e-300});
        auto spread = [](unsigned long long v) {  // 21 bits -> every third bit
            v &= 0x1fffffull;
            v = (v | v << 32) & 0x1f00000000ffffull;
            v = (v | v << 16) & 0x1f0000ff0000ffull;
            v = (v | v << 8) & 0x100f00f00f00f00full;
            v = (v | v << 4) & 0x10c30c30c30c30c3ull;
            v = (v | v << 2) & 0x1249249249249249ull;
            return v;
        };
        std::vector<std::pair<unsigned long long, int>> key(nT);
        for (size_t t = 0; t < nT; ++t) {
            unsigned long long q[3];
            for (int k = 0; k < 3; ++k)  // the same physical length per level on every axis
                q[k] = (unsigned long long)(std::min(1.0, (cen[t * 3 + size_t(k)] - lo[k]) / ext) * 524287.0);  // 19 bits
            key[t] = {spread(q[0]) | (spread(q[1]) << 1) | (spread(q[2]) << 2), int(t)};
            // coarse key: position of the tile's first run in the CALLER's order, in 16 buckets -- keeps the layout
            // monotone enough for the chunk pipeline of the host-staged entry point (a caller chunk maps to a
            // bounded range of tiles and vice versa)
            const unsigned long long bucket = nT >= 1024 ? (unsigned long long)full[t * RUNS_PER_TILE] * 16ull / (unsigned long long)nB : 0ull;
            key[t].first |= bucket << 58;  // the Morton key uses bits 0..56
        }
        std::sort(key.begin(), key.end());
        std::vector<int> ordered;
        ordered.reserve(full.size());
        for (const auto& kt : key)
            for (int r = 0; r < RUNS_PER_TILE; ++r) ordered.push_back(full[size_t(kt.second) * RUNS_PER_TILE + size_t(r)]);
        full.swap(ordered);
    }
    // tileOrderMode 2: "pencil" order.  Tiles are elongated along the caller's fastest index (a 64 x 2 x 2 brick shares
    // 128 faces with each of its four y/z neighbours but only 4 with an x neighbour), so the heavy out-of-tile gathers
    // are the ones across the tile's two SHORT directions.  Sorting the tiles lexicographically with the long
    // direction slowest puts those neighbours 1 tile and one pencil (nz/2 tiles, a few MB) away instead of one whole
    // layer of bricks (512x512 planes: 134 MB > L2); only the light x gathers (3 % of the cells) become far.
    if (tileOrderMode == 2 && full.size() >= size_t(8 * RUNS_PER_TILE)) {
        const size_t nT = full.size() / RUNS_PER_TILE;
        std::vector<double> cen(nT * 3, 0.0);
        double ext[3] = {0.0, 0.0, 0.0};
        for (size_t t = 0; t < nT; ++t) {
            double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
            for (int r = 0; r < RUNS_PER_TILE; ++r) {
                const int b = full[t * RUNS_PER_TILE + size_t(r)];
                const int c0 = b * RUN, c1 = std::min(nC, (b + 1) * RUN) - 1;
                for (int k = 0; k < 3; ++k) {
                    const double x0 = m->C[size_t(c0) * 3 + size_t(k)], x1 = m->C[size_t(c1) * 3 + size_t(k)];
                    cen[t * 3 + size_t(k)] += 0.5 * (x0 + x1) / RUNS_PER_TILE;
                    lo[k] = std::min({lo[k], x0, x1});
                    hi[k] = std::max({hi[k], x0, x1});
                }
            }
            for (int k = 0; k < 3; ++k) ext[k] += (hi[k] - lo[k]) / double(nT);
        }
        // levels along an axis: clusters of (nearly) equal centroid coordinates
        std::vector<int> level[3];
        int nLevels[3];
        for (int k = 0; k < 3; ++k) {
            std::vector<std::pair<double, int>> v(nT);
            for (size_t t = 0; t < nT; ++t) v[t] = {cen[t * 3 + size_t(k)], int(t)};
            std::sort(v.begin(), v.end());
            const double tol = 0.25 * std::max(ext[k], (v.back().first - v.front().first) / double(nT));
            level[k].assign(nT, 0);
            int lv = 0;
            double start = v[0].first;
            for (size_t i = 0; i < nT; ++i) {
                if (v[i].first - start > tol) {
                    ++lv;
                    start = v[i].first;
                }
                level[k][size_t(v[i].second)] = lv;
            }
            nLevels[k] = lv + 1;
        }
        // slowest axis = the direction across which the tiles share the FEWEST faces (the tiles' long direction),
        // judged from the face normals of the inter-tile faces, not from cell shapes
        long long crossing[3] = {0, 0, 0};
        {
            std::vector<int> tileOfRun2(size_t(nB), -1);
            for (size_t i = 0; i < full.size(); ++i) tileOfRun2[size_t(full[i])] = int(i / RUNS_PER_TILE);
            for (int f = 0; f < nIF; f += 7) {
                const int ta = tileOfRun2[size_t(m->owner[f] / RUN)], tb = tileOfRun2[size_t(m->neighbour[f] / RUN)];
                if (ta == tb) continue;
                const double sx = std::fabs(m->Sf[size_t(f) * 3]), sy = std::fabs(m->Sf[size_t(f) * 3 + 1]),
                             sz = std::fabs(m->Sf[size_t(f) * 3 + 2]);
                ++crossing[sx >= sy && sx >= sz ? 0 : (sy >= sz ? 1 : 2)];
            }
        }
        int A = 0;
        for (int k = 1; k < 3; ++k)
            if (crossing[k] < crossing[A]) A = k;
        int B = (A + 1) % 3, Cx = (A + 2) % 3;
        if (nLevels[Cx] > nLevels[B]) std::swap(B, Cx);  // fastest: the short direction with the fewest levels
        std::vector<int> idx(nT);
        for (size_t t = 0; t < nT; ++t) idx[t] = int(t);
        std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) {
            if (level[A][size_t(x)] != level[A][size_t(y)]) return level[A][size_t(x)] < level[A][size_t(y)];
            if (level[B][size_t(x)] != level[B][size_t(y)]) return level[B][size_t(x)] < level[B][size_t(y)];
            return level[Cx][size_t(x)] < level[Cx][size_t(y)];
        });
        std::vector<int> ordered;
        ordered.reserve(full.size());
        for (int t : idx)
            for (int r = 0; r < RUNS_PER_TILE; ++r) ordered.push_back(full[size_t(t) * RUNS_PER_TILE + size_t(r)]);
        full.swap(ordered);
    }
    perm.clear();
    perm.reserve(size_t(nC));
    for (const std::vector<int>* lst : {&full, &partial})
        for (int b : *lst)
            for (int c = b * RUN; c < std::min(nC, (b + 1) * RUN); ++c) perm.push_back(c);
    if (int(perm.size()) != nC) return false;  // (cannot happen; keeps a wrong layout from ever being used)
    std::vector<int> inv(static_cast<size_t>(nC), 0);
    for (int i = 0; i < nC; ++i) inv[size_t(perm[size_t(i)])] = i;
    long long remoteCompact = 0;
    for (int f = 0; f < nIF; ++f)
        if (inv[size_t(m->owner[f])] / TILE != inv[size_t(m->neighbour[f])] / TILE) remoteCompact += 2;
    remotePerCellCaller = double(remoteCaller) / nC;
    remotePerCellCompact = double(remoteCompact) / nC;
    // pays when at least one gather per cell moves from global memory into the tile
    return layoutMode == 2 || remotePerCellCaller - remotePerCellCompact >= 1.0;
}

void permuteMesh(const ssam_mesh_desc* m, PermutedMesh& pm) {
    const int nC = m->nCells, nIF = m->nInternalFaces, nF = m->nFaces;
    pm.inv.assign(size_t(nC), 0);
    for (int i = 0; i < nC; ++i) pm.inv[size_t(pm.perm[size_t(i)])] = i;
    // internal faces: stable counting sort by the new label of the (unchanged) owner
    std::vector<int> start(size_t(nC) + 1, 0);
    for (int f = 0; f < nIF; ++f) ++start[size_t(pm.inv[size_t(m->owner[f])]) + 1];
    for (int c = 0; c < nC; ++c) start[size_t(c) + 1] += start[size_t(c)];
    pm.faceKey.assign(size_t(nIF), 0);
    for (int f = 0; f < nIF; ++f) pm.faceKey[size_t(start[size_t(pm.inv[size_t(m->owner[f])])]++)] = f;
    pm.owner.resize(size_t(nF));
    pm.neighbour.resize(size_t(nIF));
    pm.Sf.resize(size_t(nF) * 3);
    pm.w.resize(size_t(nF));
    for (int g = 0; g < nF; ++g) {
        const int f = g < nIF ? pm.faceKey[size_t(g)] : g;
        pm.owner[size_t(g)] = pm.inv[size_t(m->owner[f])];
        if (g < nIF) pm.neighbour[size_t(g)] = pm.inv[size_t(m->neighbour[f])];
        for (int k = 0; k < 3; ++k) pm.Sf[size_t(g) * 3 + k] = m->Sf[size_t(f) * 3 + k];
        pm.w[size_t(g)] = m->weights[f];
    }
    pm.V.resize(size_t(nC));
    pm.C.resize(size_t(nC) * 3);
    for (int i = 0; i < nC; ++i) {
        const int c = pm.perm[size_t(i)];
        pm.V[size_t(i)] = m->V[c];
        for (int k = 0; k < 3; ++k) pm.C[size_t(i) * 3 + k] = m->C[size_t(c) * 3 + k];
    }
}

ssam_status mixtureCalcNu(ssam_ctx* ctx, const ssam_mixture_params* p) {
    TRY(need(ctx, SSAM_F_ALPHA1, "alpha1"));
    TRY(need(ctx, SSAM_F_NU1, "nu1"));
    TRY(need(ctx, SSAM_F_NU2, "nu2"));
    TRY(ensureField(ctx, SSAM_F_NU));
    CalcNuStage s{*p, ctx->plane[SSAM_F_ALPHA1][0], ctx->plane[SSAM_F_NU1][0], ctx->plane[SSAM_F_NU2][0],
                  ctx->plane[SSAM_F_NU][0]};
    LAUNCH((k_map<CalcNuStage>), mapGrid(ctx), 256, ctx->nTot, s);
    ctx->isSet[SSAM_F_NU] = true;
    return SSAM_OK;
}

}  // namespace

// ================================================================================================
extern "C" {

int ssam_abi_version(void) { return 1; }

ssam_status ssam_create(ssam_ctx** out, int device) {
    ssam_ctx* ctx = nullptr;
    CHECK(out != nullptr, SSAM_ERR_INVALID, "null out pointer");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, SSAM_ERR_CUDA,
                    std::string("no CUDA device (this path has no CPU fallback): ") + cudaGetErrorString(e));
    CHECK(device >= 0 && device < count, SSAM_ERR_INVALID, "device out of range");
    CU(cudaSetDevice(device));
    ssam_ctx* c = new ssam_ctx;
    c->device = device;
    if (const char* lm = std::getenv("SSAM_LAYOUT")) c->layoutMode = std::max(0, std::min(2, std::atoi(lm)));
    if (cudaStreamCreateWithFlags(&c->ownStream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return fail(nullptr, SSAM_ERR_CUDA, "cudaStreamCreate failed");
    }
    c->stream = c->ownStream;
    *out = c;
    return SSAM_OK;
}

ssam_status ssam_destroy(ssam_ctx* ctx) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    if (!ctx) return SSAM_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    // an in-process rank group ends with its first member: the others must not push into this context's freed slots
    for (ssam_ctx* q : ctx->localPeers) {
        if (q == ctx) continue;
        cudaSetDevice(q->device);
        cudaStreamSynchronize(q->stream);
        p2pFree(q);
        q->localPeers.clear();
        q->rank = 0;
        q->nRanks = 1;
        q->fricPatchCached = -1;
    }
    ctx->localPeers.clear();
    cudaSetDevice(ctx->device);
    if (ctx->comm) nccl().CommDestroy(ctx->comm);
    freeMesh(ctx);
    devFree(ctx->stage);
    devFree(ctx->fricSums);
    if (ctx->h2dStream) cudaStreamDestroy(ctx->h2dStream);
    if (ctx->d2hStream) cudaStreamDestroy(ctx->d2hStream);
    if (ctx->commStream) cudaStreamDestroy(ctx->commStream);
    if (ctx->evMainToComm) cudaEventDestroy(ctx->evMainToComm);
    if (ctx->evCommToMain) cudaEventDestroy(ctx->evCommToMain);
    if (ctx->evHaloIn) cudaEventDestroy(ctx->evHaloIn);
    if (ctx->evCoupledDone) cudaEventDestroy(ctx->evCoupledDone);
    if (ctx->ownStream) cudaStreamDestroy(ctx->ownStream);
    delete ctx;
    return SSAM_OK;
}

const char* ssam_last_error(const ssam_ctx* ctx) { return ctx ? ctx->err.c_str() : g_globalErr.c_str(); }

ssam_status ssam_sync(ssam_ctx* ctx) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx, SSAM_ERR_INVALID, "null context");
    CU(cudaStreamSynchronize(ctx->stream));
    return SSAM_OK;
}

ssam_status ssam_set_stream(ssam_ctx* ctx, void* s) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx, SSAM_ERR_INVALID, "null context");
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->stream = s ? reinterpret_cast<cudaStream_t>(s) : ctx->ownStream;
    return SSAM_OK;
}

ssam_status ssam_set_blocking(ssam_ctx* ctx, int blocking) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx, SSAM_ERR_INVALID, "null context");
    ctx->blocking = blocking != 0;
    return SSAM_OK;
}

int64_t ssam_launch_count(const ssam_ctx* ctx) { return ctx ? ctx->launches : 0; }

ssam_status ssam_set_host_chunks(ssam_ctx* ctx, int n) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx, SSAM_ERR_INVALID, "null context");
    ctx->hostChunks = n;
    return SSAM_OK;
}

ssam_status ssam_set_cell_kernel(ssam_ctx* ctx, int variant) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx, SSAM_ERR_INVALID, "null context");
    CHECK(variant >= 0 && variant <= 2, SSAM_ERR_INVALID,
          "cell kernel variant must be 0 (direct), 1 (tiled, one CTA per tile) or 2 (persistent, self-pipelined)");
    ctx->cellKernel = variant;
    return SSAM_OK;
}

ssam_status ssam_set_cell_layout(ssam_ctx* ctx, int mode) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx, SSAM_ERR_INVALID, "null context");
    CHECK(mode >= SSAM_LAYOUT_CALLER && mode <= SSAM_LAYOUT_COMPACT, SSAM_ERR_INVALID, "bad cell layout mode");
    ctx->layoutMode = mode;
    return SSAM_OK;
}
int ssam_cell_layout_is_compact(const ssam_ctx* ctx) { return (ctx && ctx->perm) ? 1 : 0; }

// host-only: the order ssam_mesh_set would use (no context, no GPU)
ssam_status ssam_compact_cell_order(const ssam_mesh_desc* m, int32_t* order, double* remote_caller,
                                    double* remote_compact, int* would_use) {
    if (!m || !order || m->nCells <= 0 || !m->owner || (m->nInternalFaces > 0 && !m->neighbour))
        return fail(nullptr, SSAM_ERR_INVALID, "null / empty mesh");
    for (int f = 0; f < m->nInternalFaces; ++f)
        if (m->owner[f] < 0 || m->owner[f] >= m->nCells || m->neighbour[f] < 0 || m->neighbour[f] >= m->nCells)
            return fail(nullptr, SSAM_ERR_INVALID, "owner / neighbour label out of range");
    std::vector<int> perm;
    double a = 0, b = 0;
    const bool use = computeCompactOrder(SSAM_LAYOUT_AUTO, tileOrderModeFromEnv(), m, perm, a, b);
    for (int i = 0; i < m->nCells; ++i) order[i] = perm[size_t(i)];
    if (remote_caller) *remote_caller = a;
    if (remote_compact) *remote_compact = b;
    if (would_use) *would_use = (use && m->nCells >= 2 * TILE) ? 1 : 0;
    return SSAM_OK;
}

ssam_status ssam_profile_enable(ssam_ctx* ctx, int enable) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx, SSAM_ERR_INVALID, "null context");
    ctx->profile = enable != 0;
    return SSAM_OK;
}

ssam_status ssam_profile_get(ssam_ctx* ctx, double* ms, int64_t* n) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx && ms && n, SSAM_ERR_INVALID, "null argument");
    CU(cudaStreamSynchronize(ctx->stream));
    double total = 0.0;
    for (auto& pr : ctx->profEvents) {
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, pr.first, pr.second));
        total += double(t);
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    *ms = total;
    *n = ctx->profUpdates > 0 ? ctx->profUpdates : int64_t(ctx->profEvents.size());
    ctx->profEvents.clear();
    ctx->profUpdates = 0;
    return SSAM_OK;
}

// ---- mesh -------------------------------------------------------------------------------------------
static ssam_status meshSetImpl(ssam_ctx* ctx, const ssam_mesh_desc* m) {
    CHECK(ctx && m, SSAM_ERR_INVALID, "null argument");
    CU(cudaSetDevice(ctx->device));
    CHECK(m->nCells > 0 && m->nInternalFaces >= 0 && m->nFaces >= m->nInternalFaces && m->nPatches >= 0,
          SSAM_ERR_INVALID, "bad mesh sizes");
    CHECK(m->owner && m->Sf && m->weights && m->V && m->C && m->patchStart && m->patchSize && m->patchKind &&
              m->patchNbrRank,
          SSAM_ERR_INVALID, "null mesh array");
    CU(cudaStreamSynchronize(ctx->stream));
    freeMesh(ctx);
    // tile-compact layout: permute the mesh on the host, then build everything from the permuted arrays
    PermutedMesh pm;
    ssam_mesh_desc mm;
    const ssam_mesh_desc* callerMesh = m;
    bool permuted = false;
    CHECK(m->nInternalFaces == 0 || m->neighbour, SSAM_ERR_INVALID, "null neighbour array");
    for (int f = 1; f < m->nInternalFaces; ++f)
        CHECK(m->owner[f] >= m->owner[f - 1], SSAM_ERR_INVALID,
              "owner[] of the internal faces is not ascending: the mesh is not in upper-triangular order");
    for (int f = 0; f < m->nFaces; ++f)
        CHECK(m->owner[f] >= 0 && m->owner[f] < m->nCells &&
                  (f >= m->nInternalFaces || (m->neighbour[f] >= 0 && m->neighbour[f] < m->nCells)),
              SSAM_ERR_INVALID, "owner / neighbour label out of range");
    for (int p = 0; p < m->nPatches; ++p) {
        CHECK(m->patchKind[p] == SSAM_PATCH_GENERIC || m->patchKind[p] == SSAM_PATCH_PROCESSOR ||
                  (m->patchKind[p] == SSAM_PATCH_EMPTY && m->patchSize[p] == 0),
              SSAM_ERR_INVALID,
              "unsupported patch kind: empty patches with faces (2-D / 1-D cases) carry zero-size fields in OpenFOAM and "
              "are not part of this 3-D path");
        CHECK(m->patchKind[p] != SSAM_PATCH_PROCESSOR || m->patchNbrRank[p] >= 0, SSAM_ERR_INVALID,
              "processor patch without a neighbour rank");
    }
    if (ctx->layoutMode != SSAM_LAYOUT_CALLER && m->nCells >= 2 * TILE && m->nInternalFaces > 0) {
        if (computeCompactOrder(ctx->layoutMode, tileOrderModeFromEnv(), m, pm.perm, ctx->remotePerCellCaller,
                                ctx->remotePerCellCompact)) {
            permuteMesh(m, pm);
            mm = *m;
            mm.owner = pm.owner.data();
            mm.neighbour = pm.neighbour.data();
            mm.Sf = pm.Sf.data();
            mm.weights = pm.w.data();
            mm.V = pm.V.data();
            mm.C = pm.C.data();
            m = &mm;
            permuted = true;
        }
    }
    ctx->nCells = m->nCells;
    ctx->nIF = m->nInternalFaces;
    ctx->nFaces = m->nFaces;
    ctx->nBnd = m->nFaces - m->nInternalFaces;
    ctx->nTot = ctx->nCells + ctx->nBnd;
    ctx->nPatches = m->nPatches;
    ctx->patchStart.assign(m->patchStart, m->patchStart + m->nPatches);
    ctx->patchSize.assign(m->patchSize, m->patchSize + m->nPatches);
    ctx->patchKind.assign(m->patchKind, m->patchKind + m->nPatches);
    ctx->patchNbrRank.assign(m->patchNbrRank, m->patchNbrRank + m->nPatches);
    ctx->patchUbc.assign(m->nPatches, SSAM_BC_VALUE);
    // patches must tile [nIF, nFaces) in order
    int expect = ctx->nIF;
    for (int p = 0; p < m->nPatches; ++p) {
        CHECK(ctx->patchStart[p] == expect && ctx->patchSize[p] >= 0, SSAM_ERR_INVALID,
              "patches must be contiguous and ordered after the internal faces");
        expect += ctx->patchSize[p];
    }
    CHECK(expect == ctx->nFaces, SSAM_ERR_INVALID, "patches do not cover the boundary faces");

    const size_t nF = size_t(ctx->nFaces), nC = size_t(ctx->nCells), nB = size_t(ctx->nBnd);
    TRY(devAlloc(ctx, &ctx->owner, nF));
    TRY(devAlloc(ctx, &ctx->neighbour, size_t(ctx->nIF) + 8));  // padded: bulk copies round to 16 bytes
    CU(cudaMemcpyAsync(ctx->owner, m->owner, nF * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->nIF)
        CU(cudaMemcpyAsync(ctx->neighbour, m->neighbour, size_t(ctx->nIF) * sizeof(int), cudaMemcpyHostToDevice,
                           ctx->stream));
    // {Sf, w} -> one 32-byte record per face
    TRY(devAlloc(ctx, &ctx->geo, nF));
    {
        double *dSf = nullptr, *dW = nullptr;
        TRY(devAlloc(ctx, &dSf, 3 * nF));
        TRY(devAlloc(ctx, &dW, nF));
        CU(cudaMemcpyAsync(dSf, m->Sf, 3 * nF * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(dW, m->weights, nF * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        LAUNCH(k_pack_geo, gridFor((long long)nF, 256), 256, dSf, dW, ctx->nFaces, ctx->geo);
        CU(cudaStreamSynchronize(ctx->stream));
        cudaFree(dSf);
        cudaFree(dW);
    }
    TRY(devAlloc(ctx, &ctx->V, nC + 4));
    CU(cudaMemcpyAsync(ctx->V, m->V, nC * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ctx->sumV = 0.0;
    for (size_t i = 0; i < nC; ++i) ctx->sumV += callerMesh->V[i];  // Foam::sum order (the caller's cell order)
    TRY(devAlloc(ctx, &ctx->magSf_b, nB));
    TRY(devAlloc(ctx, &ctx->deltaCoeffs_b, nB));
    if (nB) {
        CHECK(m->Cf_b && m->magSf_b && m->deltaCoeffs_b, SSAM_ERR_INVALID, "null boundary geometry");
        ctx->hostCfB.assign(m->Cf_b, m->Cf_b + 3 * nB);
        ctx->hostMagSfB.assign(m->magSf_b, m->magSf_b + nB);
        CU(cudaMemcpyAsync(ctx->magSf_b, m->magSf_b, nB * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->deltaCoeffs_b, m->deltaCoeffs_b, nB * sizeof(double), cudaMemcpyHostToDevice,
                           ctx->stream));
    }
    for (int k = 0; k < 3; ++k) TRY(devAlloc(ctx, &ctx->Cpl[k], size_t(ctx->nTot) + 2));
    TRY(uploadAoS(ctx, m->C, ctx->nCells, 3, ctx->Cpl, 0));
    CU(cudaStreamSynchronize(ctx->stream));
    if (nB) TRY(uploadAoS(ctx, m->Cf_b, ctx->nBnd, 3, ctx->Cpl, ctx->nCells));
    for (int k = 0; k < 9; ++k) TRY(devAlloc(ctx, &ctx->gradB[k], nB));
    rebuildBndFlags(ctx);
    TRY(devAlloc(ctx, &ctx->bndFlags, nB));
    if (nB)
        CU(cudaMemcpyAsync(ctx->bndFlags, ctx->bndFlagsHost.data(), nB, cudaMemcpyHostToDevice, ctx->stream));
    // halo maps: faceCells of every processor patch in patch-face order
    ctx->haloPatch.clear();
    ctx->haloOffsets.assign(1, 0);
    for (int p = 0; p < ctx->nPatches; ++p)
        if (ctx->patchKind[p] == SSAM_PATCH_PROCESSOR) {
            ctx->haloPatch.push_back(p);
            ctx->haloOffsets.push_back(ctx->haloOffsets.back() + ctx->patchSize[p]);
        }
    ctx->nProcPatches = int(ctx->haloPatch.size());
    ctx->nProcFaces = ctx->haloOffsets.back();
    CHECK(ctx->nProcPatches <= 32, SSAM_ERR_INVALID, "more than 32 processor patches on one rank");
    TRY(devAlloc(ctx, &ctx->haloSendCells, size_t(ctx->nProcFaces)));
    for (int p = 0; p < ctx->nProcPatches; ++p) {
        const int patch = ctx->haloPatch[p];
        CU(cudaMemcpyAsync(ctx->haloSendCells + ctx->haloOffsets[p], ctx->owner + ctx->patchStart[patch],
                           size_t(ctx->patchSize[patch]) * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    if (permuted) {
        TRY(devAlloc(ctx, &ctx->faceKey, size_t(ctx->nIF)));
        CU(cudaMemcpyAsync(ctx->faceKey, pm.faceKey.data(), size_t(ctx->nIF) * sizeof(int), cudaMemcpyHostToDevice,
                           ctx->stream));
        TRY(devAlloc(ctx, &ctx->origOwner, nF));
        TRY(devAlloc(ctx, &ctx->origNeighbour, size_t(ctx->nIF) + 8));
        CU(cudaMemcpyAsync(ctx->origOwner, callerMesh->owner, nF * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->origNeighbour, callerMesh->neighbour, size_t(ctx->nIF) * sizeof(int),
                           cudaMemcpyHostToDevice, ctx->stream));
    }
    TRY(buildAddressing(ctx));
    {
        std::vector<int> dist;
        dist.reserve(size_t(ctx->nIF) / 8 + 1);
        for (int f = 0; f < ctx->nIF; f += 8) dist.push_back(std::abs(m->neighbour[f] / TILE - m->owner[f] / TILE));
        ctx->pfBandwidthTiles = ctx->bandwidthTiles;
        if (!dist.empty()) {
            const size_t k = std::min(dist.size() - 1, size_t(0.97 * double(dist.size())));
            std::nth_element(dist.begin(), dist.begin() + long(k), dist.end());
            ctx->pfBandwidthTiles = std::min(ctx->bandwidthTiles, dist[k] + 1);
        }
    }
    if (ctx->nProcPatches == 0) TRY(buildTileOrder(ctx, m));
    if (permuted) {
        // from here on uploads / downloads of cell values go through the permutation
        TRY(devAlloc(ctx, &ctx->perm, nC));
        CU(cudaMemcpyAsync(ctx->perm, pm.perm.data(), nC * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        TRY(devAlloc(ctx, &ctx->invPerm, nC));
        CU(cudaMemcpyAsync(ctx->invPerm, pm.inv.data(), nC * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->permHost.swap(pm.perm);
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return SSAM_OK;
}

// the mesh becomes usable only when the whole build succeeded; a failed build leaves no half-built mesh behind
ssam_status ssam_mesh_set(ssam_ctx* ctx, const ssam_mesh_desc* m) {
    if (ctx) cudaSetDevice(ctx->device);
    const ssam_status st = meshSetImpl(ctx, m);
    if (ctx) {
        if (st == SSAM_OK) {
            ctx->haveMesh = true;
        } else {
            const std::string e = ctx->err;
            cudaStreamSynchronize(ctx->stream);
            cudaGetLastError();
            freeMesh(ctx);
            ctx->err = e;
        }
    }
    return st;
}

ssam_status ssam_patch_u_bc_set(ssam_ctx* ctx, const int32_t* kinds) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx && ctx->haveMesh && kinds, SSAM_ERR_INVALID, "mesh not set / null kinds");
    for (int p = 0; p < ctx->nPatches; ++p) {
        CHECK(kinds[p] == SSAM_BC_VALUE || kinds[p] == SSAM_BC_ZERO_GRADIENT, SSAM_ERR_INVALID, "bad U bc kind");
        ctx->patchUbc[p] = kinds[p];
    }
    rebuildBndFlags(ctx);
    if (ctx->nBnd)
        CU(cudaMemcpyAsync(ctx->bndFlags, ctx->bndFlagsHost.data(), size_t(ctx->nBnd), cudaMemcpyHostToDevice,
                           ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SSAM_OK;
}

// ---- addressing getters -------------------------------------------------------------------------------
static ssam_status ensureCellFaces(ssam_ctx* ctx) {
    if (ctx->cellFaces) return SSAM_OK;
    TRY(devAlloc(ctx, &ctx->cellFacesStart, size_t(ctx->nCells) + 1));
    TRY(devAlloc(ctx, &ctx->cellFaces, size_t(2) * ctx->nIF + ctx->nBnd));
    LAUNCH(k_cell_faces, gridFor(ctx->nCells + 1, 256), 256, ctx->ownerStart, ctx->extraStart, ctx->extra, ctx->nCells,
           ctx->nIF, ctx->cellFacesStart, ctx->cellFaces);
    return SSAM_OK;
}

static ssam_status ensureColouring(ssam_ctx* ctx) {
    if (ctx->faceColour) return SSAM_OK;
    TRY(devAlloc(ctx, &ctx->faceColour, size_t(ctx->nIF)));
    CU(cudaMemsetAsync(ctx->faceColour, 0xff, size_t(ctx->nIF) * sizeof(int), ctx->stream));
    int* rem = nullptr;
    TRY(devAlloc(ctx, &rem, 1));
    for (int sweep = 0; sweep < 1 << 20; ++sweep) {
        CU(cudaMemsetAsync(rem, 0, sizeof(int), ctx->stream));
        LAUNCH(k_colour_sweep, gridFor(ctx->nIF, 256), 256, ctx->owner, ctx->neighbour, ctx->ownerStart,
               ctx->extraStart, ctx->extra, ctx->nIF, ctx->faceColour, rem);
        int h = 0;
        CU(cudaMemcpyAsync(&h, rem, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (h == 0) break;
    }
    cudaFree(rem);
    return SSAM_OK;
}

// Tile-compact layout: the tables the getters expose are defined on the caller's numbering, so they are built
// (on demand, by the same kernels) from the caller's owner/neighbour kept on the device, in a private context.
static ssam_status ensureRefAddr(ssam_ctx* ctx) {
    if (ctx->refAddr) return SSAM_OK;
    ssam_ctx* r = new ssam_ctx;
    r->device = ctx->device;
    r->stream = ctx->stream;
    r->blocking = true;
    r->layoutMode = SSAM_LAYOUT_CALLER;
    r->nCells = ctx->nCells;
    r->nIF = ctx->nIF;
    r->nFaces = ctx->nFaces;
    r->nBnd = ctx->nBnd;
    r->nTot = ctx->nTot;
    r->nPatches = ctx->nPatches;
    r->patchStart = ctx->patchStart;
    r->patchSize = ctx->patchSize;
    r->patchKind = ctx->patchKind;
    r->patchNbrRank = ctx->patchNbrRank;
    r->patchUbc = ctx->patchUbc;
    r->haloPatch = ctx->haloPatch;
    r->haloOffsets = ctx->haloOffsets;
    r->nProcPatches = ctx->nProcPatches;
    r->nProcFaces = ctx->nProcFaces;
    r->haveMesh = true;
    ctx->refAddr = r;
    auto bail = [&](ssam_status st) {
        ctx->err = r->err;
        freeMesh(r);
        delete r;
        ctx->refAddr = nullptr;
        return st;
    };
#define REF_TRY(x)                                  \
    do {                                            \
        ssam_status st_ = (x);                      \
        if (st_ != SSAM_OK) return bail(st_);       \
    } while (0)
    REF_TRY(devAlloc(r, &r->owner, size_t(r->nFaces)));
    REF_TRY(devAlloc(r, &r->neighbour, size_t(r->nIF) + 8));
    REF_TRY(devAlloc(r, &r->bndFlags, size_t(r->nBnd)));
    REF_TRY(devAlloc(r, &r->haloSendCells, size_t(r->nProcFaces)));
    cudaMemcpyAsync(r->owner, ctx->origOwner, size_t(r->nFaces) * sizeof(int), cudaMemcpyDeviceToDevice, r->stream);
    cudaMemcpyAsync(r->neighbour, ctx->origNeighbour, size_t(r->nIF) * sizeof(int), cudaMemcpyDeviceToDevice, r->stream);
    if (r->nBnd)
        cudaMemcpyAsync(r->bndFlags, ctx->bndFlags, size_t(r->nBnd), cudaMemcpyDeviceToDevice, r->stream);
    for (int p = 0; p < r->nProcPatches; ++p) {
        const int patch = r->haloPatch[size_t(p)];
        cudaMemcpyAsync(r->haloSendCells + r->haloOffsets[size_t(p)], r->owner + r->patchStart[size_t(patch)],
                        size_t(r->patchSize[size_t(patch)]) * sizeof(int), cudaMemcpyDeviceToDevice, r->stream);
    }
    REF_TRY(buildAddressing(r));
#undef REF_TRY
    return SSAM_OK;
}

ssam_status ssam_addressing_size(ssam_ctx* ctx, int which, int64_t* n) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx && ctx->haveMesh && n, SSAM_ERR_INVALID, "mesh not set");
    if (which == SSAM_ADDR_CELL_ORDER) {
        *n = ctx->nCells;
        return SSAM_OK;
    }
    if (which == SSAM_ADDR_FACE_ORDER) {
        *n = ctx->nIF;
        return SSAM_OK;
    }
    switch (which) {
        case SSAM_ADDR_OWNER_START:
        case SSAM_ADDR_EXTRA_START:
        case SSAM_ADDR_CELL_FACES_START: *n = int64_t(ctx->nCells) + 1; break;
        case SSAM_ADDR_EXTRA_FACE:
        case SSAM_ADDR_EXTRA_OTHER: *n = ctx->nExtra; break;
        case SSAM_ADDR_CELL_FACES: *n = int64_t(2) * ctx->nIF + ctx->nBnd; break;
        case SSAM_ADDR_HALO_SEND_CELLS: *n = ctx->nProcFaces; break;
        case SSAM_ADDR_HALO_PATCH_OFFSETS: *n = ctx->nProcPatches + 1; break;
        case SSAM_ADDR_FRIC_CELLS: *n = ctx->fricPatchCached >= 0 ? ctx->patchSize[ctx->fricPatchCached] : 0; break;
        case SSAM_ADDR_FACE_COLOUR: *n = ctx->nIF; break;
        default: return fail(ctx, SSAM_ERR_INVALID, "bad addressing id");
    }
    return SSAM_OK;
}

ssam_status ssam_addressing_get(ssam_ctx* ctx, int which, int32_t* out, int64_t n) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    int64_t have = 0;
    TRY(ssam_addressing_size(ctx, which, &have));
    CHECK(out && n == have, SSAM_ERR_INVALID, "size mismatch");
    if (n == 0) return SSAM_OK;
    if (which == SSAM_ADDR_CELL_ORDER) {
        for (int64_t i = 0; i < n; ++i) out[i] = ctx->permHost.empty() ? int32_t(i) : ctx->permHost[size_t(i)];
        return SSAM_OK;
    }
    if (which == SSAM_ADDR_FACE_ORDER) {
        if (ctx->faceKey) {
            CU(cudaMemcpyAsync(out, ctx->faceKey, size_t(n) * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
        } else {
            for (int64_t i = 0; i < n; ++i) out[i] = int32_t(i);
        }
        return SSAM_OK;
    }
    if (ctx->perm) {  // tables in the caller's numbering
        TRY(ensureRefAddr(ctx));
        ctx->refAddr->fricPatchCached = ctx->fricPatchCached;
        const ssam_status st = ssam_addressing_get(ctx->refAddr, which, out, n);
        if (st != SSAM_OK) ctx->err = ctx->refAddr->err;
        return st;
    }
    const int* src = nullptr;
    int* tmp = nullptr;
    switch (which) {
        case SSAM_ADDR_OWNER_START: src = ctx->ownerStart; break;
        case SSAM_ADDR_EXTRA_START: src = ctx->extraStart; break;
        case SSAM_ADDR_EXTRA_FACE:
        case SSAM_ADDR_EXTRA_OTHER: {
            TRY(devAlloc(ctx, &tmp, size_t(2) * n));
            LAUNCH(k_split_int2, gridFor(n, 256), 256, ctx->extra, int(n), tmp, tmp + n);
            src = which == SSAM_ADDR_EXTRA_FACE ? tmp : tmp + n;
            break;
        }
        case SSAM_ADDR_CELL_FACES_START:
            TRY(ensureCellFaces(ctx));
            src = ctx->cellFacesStart;
            break;
        case SSAM_ADDR_CELL_FACES:
            TRY(ensureCellFaces(ctx));
            src = ctx->cellFaces;
            break;
        case SSAM_ADDR_HALO_SEND_CELLS: src = ctx->haloSendCells; break;
        case SSAM_ADDR_HALO_PATCH_OFFSETS:
            std::memcpy(out, ctx->haloOffsets.data(), size_t(n) * sizeof(int));
            return SSAM_OK;
        case SSAM_ADDR_FRIC_CELLS: src = ctx->owner + ctx->patchStart[ctx->fricPatchCached]; break;
        case SSAM_ADDR_FACE_COLOUR:
            TRY(ensureColouring(ctx));
            src = ctx->faceColour;
            break;
    }
    CU(cudaMemcpyAsync(out, src, size_t(n) * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (tmp) cudaFree(tmp);
    return SSAM_OK;
}

// ---- fields -------------------------------------------------------------------------------------------
ssam_status ssam_field_upload(ssam_ctx* ctx, int field, const double* internal, const double* boundary) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx, SSAM_ERR_INVALID, "null context");
    TRY(ensureField(ctx, field));
    const int nc = ncompOf(field);
    if (internal) {
        TRY(uploadAoS(ctx, internal, ctx->nCells, nc, ctx->plane[field], 0));
        if (nc > 1) CU(cudaStreamSynchronize(ctx->stream));  // staging buffer reuse
    }
    if (boundary) TRY(uploadAoS(ctx, boundary, ctx->nBnd, nc, ctx->plane[field], ctx->nCells));
    ctx->isSet[field] = true;
    if (field == SSAM_F_NU2) ctx->nu2Filled = false;
    CU(cudaStreamSynchronize(ctx->stream));
    return SSAM_OK;
}

ssam_status ssam_field_download(ssam_ctx* ctx, int field, double* internal, double* boundary) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx, SSAM_ERR_INVALID, "null context");
    CHECK(field >= 0 && field < SSAM_F_COUNT, SSAM_ERR_INVALID, "bad field id");
    TRY(need(ctx, field, "download of a field that was never uploaded or computed"));
    const int nc = ncompOf(field);
    if (internal) TRY(downloadAoS(ctx, internal, ctx->nCells, nc, ctx->plane[field], 0));
    if (boundary) TRY(downloadAoS(ctx, boundary, ctx->nBnd, nc, ctx->plane[field], ctx->nCells));
    CU(cudaStreamSynchronize(ctx->stream));
    return SSAM_OK;
}

ssam_status ssam_field_device_ptr(ssam_ctx* ctx, int field, int comp, void** ptr, int64_t* len) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx && ptr, SSAM_ERR_INVALID, "null argument");
    TRY(ensureField(ctx, field));
    CHECK(comp >= 0 && comp < ncompOf(field), SSAM_ERR_INVALID, "bad component");
    *ptr = ctx->plane[field][comp];
    if (len) *len = ctx->nTot;
    ctx->isSet[field] = true;  // the caller fills it on the device
    if (field == SSAM_F_NU2) ctx->nu2Filled = false;
    CU(cudaStreamSynchronize(ctx->stream));
    return SSAM_OK;
}

int ssam_field_is_set(const ssam_ctx* ctx, int field) {
    return (ctx && field >= 0 && field < SSAM_F_COUNT && ctx->isSet[field]) ? 1 : 0;
}

// ---- boundary condition ---------------------------------------------------------------------------------
ssam_status ssam_bc_rotational_slip(ssam_ctx* ctx, const ssam_rotslip_params* p) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx && p && ctx->haveMesh, SSAM_ERR_INVALID, "mesh not set / null params");
    CHECK(p->patch >= 0 && p->patch < ctx->nPatches, SSAM_ERR_INVALID, "patch out of range");
    CHECK(p->model == SSAM_ROTSLIP_CONSTANT || p->model == SSAM_ROTSLIP_EXPONENTIAL, SSAM_ERR_UNKNOWN_MODEL,
          "Unknown rotational slip patch field type");
    TRY(ensureField(ctx, SSAM_F_U));
    const int n = ctx->patchSize[p->patch];
    LAUNCH(k_bc_rotslip, gridFor(n, 128), 128, *p, ctx->patchStart[p->patch] - ctx->nIF, n, ctx->nCells, ctx->Cpl[0],
           ctx->Cpl[1], ctx->Cpl[2], ctx->plane[SSAM_F_U][0], ctx->plane[SSAM_F_U][1], ctx->plane[SSAM_F_U][2]);
    return finish(ctx);
}

// ---- staged entry points ----------------------------------------------------------------------------------
ssam_status ssam_grad_u(ssam_ctx* ctx) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx, SSAM_ERR_INVALID, "null context");
    TRY(gradOrStrain(ctx, false, 0.0, SSAM_F_GRADU));
    return finish(ctx);
}

ssam_status ssam_strain_rate(ssam_ctx* ctx, double factor, int out_field) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx, SSAM_ERR_INVALID, "null context");
    CHECK(out_field == SSAM_F_EPSDOT || out_field == SSAM_F_STRAINRATE, SSAM_ERR_INVALID,
          "strain rate goes to EPSDOT or STRAINRATE");
    TRY(gradOrStrain(ctx, true, factor, out_field));
    return finish(ctx);
}

ssam_status ssam_viscosity_correct(ssam_ctx* ctx, int phase, const ssam_viscosity_params* p) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx, SSAM_ERR_INVALID, "null context");
    TRY(viscosityCorrect(ctx, phase, p));
    return finish(ctx);
}

ssam_status ssam_mixture_calc_nu(ssam_ctx* ctx, const ssam_mixture_params* p) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx && p, SSAM_ERR_INVALID, "null argument");
    TRY(mixtureCalcNu(ctx, p));
    return finish(ctx);
}

ssam_status ssam_thermo_correct(ssam_ctx* ctx, const ssam_viscosity_params* v1, const ssam_viscosity_params* v2,
                                const ssam_mixture_params* mix) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx && mix, SSAM_ERR_INVALID, "null argument");
    TRY(viscosityCorrect(ctx, 1, v1));  // nuModel1_->correct()   (...Mixture.C:43)
    TRY(viscosityCorrect(ctx, 2, v2));  // nuModel2_->correct()   (:44)
    TRY(mixtureCalcNu(ctx, mix));       // :46-53
    return finish(ctx);
}

ssam_status ssam_mixture_mu(ssam_ctx* ctx, const ssam_mixture_params* p) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx && p, SSAM_ERR_INVALID, "null argument");
    TRY(need(ctx, SSAM_F_ALPHA1, "alpha1"));
    TRY(need(ctx, SSAM_F_NU1, "nu1"));
    TRY(need(ctx, SSAM_F_NU2, "nu2"));
    TRY(ensureField(ctx, SSAM_F_MUEFF));
    MuStage s{*p, ctx->plane[SSAM_F_ALPHA1][0], ctx->plane[SSAM_F_NU1][0], ctx->plane[SSAM_F_NU2][0],
              ctx->plane[SSAM_F_MUEFF][0]};
    LAUNCH((k_map<MuStage>), mapGrid(ctx), 256, ctx->nTot, s);
    ctx->isSet[SSAM_F_MUEFF] = true;
    return finish(ctx);
}

static ssam_status mixtureLinear(ssam_ctx* ctx, int outField, double v1, double v2) {
    TRY(need(ctx, SSAM_F_ALPHA1, "alpha1"));
    TRY(ensureField(ctx, outField));
    LinearStage s{v1, v2, ctx->plane[SSAM_F_ALPHA1][0], ctx->plane[outField][0]};
    LAUNCH((k_map<LinearStage>), mapGrid(ctx), 256, ctx->nTot, s);
    ctx->isSet[outField] = true;
    return finish(ctx);
}

ssam_status ssam_mixture_rho(ssam_ctx* ctx, const ssam_mixture_params* p) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx && p, SSAM_ERR_INVALID, "null argument");
    return mixtureLinear(ctx, SSAM_F_RHO, p->rho1, p->rho2);
}

ssam_status ssam_mixture_cp(ssam_ctx* ctx, const ssam_mixture_params* p) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx && p, SSAM_ERR_INVALID, "null argument");
    return mixtureLinear(ctx, SSAM_F_CP, p->cp1, p->cp2);
}

ssam_status ssam_mixture_k(ssam_ctx* ctx, const ssam_mixture_params* p) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx && p, SSAM_ERR_INVALID, "null argument");
    TRY(need(ctx, SSAM_F_ALPHA1, "alpha1"));
    TRY(need(ctx, SSAM_F_T, "temp"));
    double s1, s2;
    TRY(unitsShift(ctx, p->k1, &s1));
    TRY(unitsShift(ctx, p->k2, &s2));
    TRY(ensureField(ctx, SSAM_F_K));
    KStage s{*p, s1, s2, ctx->plane[SSAM_F_ALPHA1][0], ctx->plane[SSAM_F_T][0], ctx->plane[SSAM_F_K][0]};
    LAUNCH((k_map<KStage>), mapGrid(ctx), 256, ctx->nTot, s);
    ctx->isSet[SSAM_F_K] = true;
    return finish(ctx);
}

ssam_status ssam_mixture_k_phase(ssam_ctx* ctx, const ssam_mixture_params* p, int phase) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx && p && ctx->haveMesh, SSAM_ERR_INVALID, "mesh not set / null params");
    CHECK(phase == 1 || phase == 2, SSAM_ERR_INVALID, "phase must be 1 or 2");
    TRY(need(ctx, SSAM_F_T, "temp"));
    const int out = phase == 1 ? SSAM_F_K1 : SSAM_F_K2;
    TRY(ensureField(ctx, out));
    KPhaseStage st;
    st.kc = phase == 1 ? p->k1 : p->k2;
    TRY(unitsShift(ctx, st.kc, &st.shift));
    st.T = ctx->plane[SSAM_F_T][0];
    st.k = ctx->plane[out][0];
    LAUNCH((k_map<KPhaseStage>), mapGrid(ctx), 256, ctx->nTot, st);
    ctx->isSet[out] = true;
    return finish(ctx);
}

ssam_status ssam_solver_rho_rhocp(ssam_ctx* ctx, const ssam_mixture_params* p) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx && p, SSAM_ERR_INVALID, "null argument");
    TRY(need(ctx, SSAM_F_ALPHA1, "alpha1"));
    TRY(ensureField(ctx, SSAM_F_RHO_SOLVER));
    TRY(ensureField(ctx, SSAM_F_RHOCP));
    SolverRhoStage s{*p, ctx->plane[SSAM_F_ALPHA1][0], ctx->plane[SSAM_F_RHO_SOLVER][0], ctx->plane[SSAM_F_RHOCP][0]};
    LAUNCH((k_map<SolverRhoStage>), mapGrid(ctx), 256, ctx->nTot, s);
    ctx->isSet[SSAM_F_RHO_SOLVER] = ctx->isSet[SSAM_F_RHOCP] = true;
    return finish(ctx);
}

ssam_status ssam_stress_strain(ssam_ctx* ctx, const ssam_stress_strain_params* p) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx && p, SSAM_ERR_INVALID, "null argument");
    TRY(need(ctx, SSAM_F_EPSDOT, "epsilonDotEq"));
    TRY(need(ctx, SSAM_F_T, "temp"));
    TRY(need(ctx, SSAM_F_MUEFF, "muEff"));
    TRY(need(ctx, SSAM_F_PRGH, "p_rgh"));
    TRY(gradOrStrain(ctx, false, 0.0, SSAM_F_GRADU));  // volTensorField gradU = fvc::grad(U)  (:17)
    const int outs[] = {SSAM_F_Z, SSAM_F_SIGMAF_PP, SSAM_F_DEPSEQ, SSAM_F_EPSEQ, SSAM_F_SIGMAVM};
    const bool accumulate = ctx->isSet[SSAM_F_EPSEQ];
    for (int f : outs) TRY(ensureField(ctx, f));
    StressStrainStage s;
    s.p = *p;
    s.logA = std::log(p->A);
    s.invN = 1 / p->n;
    for (int k = 0; k < 9; ++k) s.g[k] = ctx->plane[SSAM_F_GRADU][k];
    s.eps = ctx->plane[SSAM_F_EPSDOT][0];
    s.T = ctx->plane[SSAM_F_T][0];
    s.mu = ctx->plane[SSAM_F_MUEFF][0];
    s.prgh = ctx->plane[SSAM_F_PRGH][0];
    s.Z = ctx->plane[SSAM_F_Z][0];
    s.sigmaf = ctx->plane[SSAM_F_SIGMAF_PP][0];
    s.dEq = ctx->plane[SSAM_F_DEPSEQ][0];
    s.epsEq = ctx->plane[SSAM_F_EPSEQ][0];
    s.sigmavm = ctx->plane[SSAM_F_SIGMAVM][0];
    s.accumulate = accumulate ? 1 : 0;
    LAUNCH((k_map<StressStrainStage>), mapGrid(ctx), 256, ctx->nTot, s);
    for (int f : outs) ctx->isSet[f] = true;
    return finish(ctx);
}

// explicit part of turbulence.divDevRhoReff(rho, U)  (fluid/UEqn.H:13)
ssam_status ssam_div_dev_rho_reff_explicit(ssam_ctx* ctx) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx, SSAM_ERR_INVALID, "null context");
    TRY(need(ctx, SSAM_F_GRADU, "gradU (ssam_grad_u, or ssam_update_fused with flag 1)"));
    TRY(need(ctx, SSAM_F_RHO_SOLVER, "rho (ssam_solver_rho_rhocp)"));
    TRY(need(ctx, SSAM_F_NU, "nu (thermo.correct())"));
    TRY(ensureField(ctx, SSAM_F_DIVDEVRHOREFF));
    if (ctx->nProcPatches) {  // boundary values of the operands on processor patches = neighbour cell values
        const int ops[3] = {SSAM_F_GRADU, SSAM_F_RHO_SOLVER, SSAM_F_NU};
        TRY(haloExchangeFields(ctx, ops, 3, ctx->stream));
    }
    DivDevArgs a;
    a.nCells = ctx->nCells;
    a.nIF = ctx->nIF;
    a.nBnd = ctx->nBnd;
    a.geo = ctx->geo;
    a.neighbour = ctx->neighbour;
    a.ownerStart = ctx->ownerStart;
    a.extraStart = ctx->extraStart;
    a.extra = ctx->extra;
    a.V = ctx->V;
    for (int k = 0; k < 9; ++k) a.g[k] = ctx->plane[SSAM_F_GRADU][k];
    a.rho = ctx->plane[SSAM_F_RHO_SOLVER][0];
    a.nu = ctx->plane[SSAM_F_NU][0];
    for (int k = 0; k < 3; ++k) a.out[k] = ctx->plane[SSAM_F_DIVDEVRHOREFF][k];
    LAUNCH(k_div_dev_cells, gridFor(ctx->nCells, 128), 128, a);
    if (ctx->nProcPatches) TRY(haloExchangeField(ctx, SSAM_F_DIVDEVRHOREFF));
    ctx->isSet[SSAM_F_DIVDEVRHOREFF] = true;
    return finish(ctx);
}

// interfaceProperties::correct() -> calculateK()  (immiscibleIncompressibleTwoPhaseThermalMixture.H:80)
ssam_status ssam_interface_correct(ssam_ctx* ctx, double deltaN, unsigned flags) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx, SSAM_ERR_INVALID, "null context");
    TRY(need(ctx, SSAM_F_ALPHA1, "alpha1"));
    TRY(ensureField(ctx, SSAM_F_GRADALPHA));
    TRY(ensureField(ctx, SSAM_F_CURVATURE));
    TRY(ensureFaceField(ctx, SSAM_FF_NHATF));
    IfaceArgs a;
    std::memset(&a, 0, sizeof(a));
    a.nCells = ctx->nCells;
    a.nIF = ctx->nIF;
    a.nBnd = ctx->nBnd;
    a.geo = ctx->geo;
    a.owner = ctx->owner;
    a.neighbour = ctx->neighbour;
    a.ownerStart = ctx->ownerStart;
    a.extraStart = ctx->extraStart;
    a.extra = ctx->extra;
    a.V = ctx->V;
    a.magSf_b = ctx->magSf_b;
    a.deltaCoeffs_b = ctx->deltaCoeffs_b;
    a.bndFlags = ctx->bndFlags;
    a.alpha = ctx->plane[SSAM_F_ALPHA1][0];
    for (int k = 0; k < 3; ++k) a.g[k] = ctx->plane[SSAM_F_GRADALPHA][k];
    a.nHatf = ctx->faceField[SSAM_FF_NHATF];
    a.K = ctx->plane[SSAM_F_CURVATURE][0];
    a.deltaN = deltaN;
    const int gc = gridFor(ctx->nCells, 256), gb = gridFor(ctx->nBnd, 128), gf = gridFor(ctx->nFaces, 256);
    if (ctx->nProcPatches && (flags & 2)) TRY(haloExchangeField(ctx, SSAM_F_ALPHA1));
    const bool tiled = ctx->cellKernel >= 1 && ctx->nTilesOversize == 0;
    a.tileDesc = ctx->tileDesc;
    a.tileCount = ctx->nTiles;
    a.pfDist = ctx->pfDist;
    a.pfDistU = a.pfDist + ctx->pfBandwidthTiles;
    const int smem = int(sizeof(TileSmem));
    if (tiled) {
        static bool configured[64] = {false};
        if (!configured[ctx->device & 63]) {
            CU(cudaFuncSetAttribute(k_iface_tiled<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            CU(cudaFuncSetAttribute(k_iface_tiled<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            configured[ctx->device & 63] = true;
        }
        k_iface_tiled<0><<<ctx->nTiles, TILE, smem, ctx->stream>>>(a);
        ctx->launches++;
        CU(cudaGetLastError());
    } else {
        LAUNCH(k_iface_grad_cells, gc, 256, a);
    }
    if (ctx->nProcPatches) TRY(haloExchangeField(ctx, SSAM_F_GRADALPHA));
    if (ctx->nBnd) LAUNCH(k_iface_grad_bnd, gb, 128, a);
    if (tiled) {
        k_iface_tiled<1><<<ctx->nTiles, TILE, smem, ctx->stream>>>(a);
        ctx->launches++;
        CU(cudaGetLastError());
    } else {
        LAUNCH(k_iface_nhatf, gf, 256, a);
        LAUNCH(k_iface_div_cells, gc, 256, a);
    }
    if (ctx->nProcPatches) TRY(haloExchangeField(ctx, SSAM_F_CURVATURE));
    ctx->isSet[SSAM_F_GRADALPHA] = ctx->isSet[SSAM_F_CURVATURE] = true;
    ctx->faceFieldSet[SSAM_FF_NHATF] = true;
    return finish(ctx);
}

ssam_status ssam_face_field_download(ssam_ctx* ctx, int ff, double* internal_faces, double* boundary_faces) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx, SSAM_ERR_INVALID, "null context");
    CHECK(ff >= 0 && ff < SSAM_FF_COUNT, SSAM_ERR_INVALID, "bad face field id");
    CHECK(ctx->faceField[ff] && ctx->faceFieldSet[ff], SSAM_ERR_MISSING_FIELD,
          "face field not available (nHatf: ssam_interface_correct, muf/nuf: ssam_mixture_muf/nuf)");
    const double* src = ctx->faceField[ff];
    if (internal_faces && ctx->nIF) {
        const double* from = src;
        if (ctx->faceKey) {  // tile-compact layout: the device holds the internal faces in its own order
            TRY(ensureStage(ctx, size_t(ctx->nIF)));
            LAUNCH(k_scatter_by_key, gridFor(ctx->nIF, 256), 256, src, ctx->faceKey, ctx->nIF, ctx->stage);
            from = ctx->stage;
        }
        CU(cudaMemcpyAsync(internal_faces, from, size_t(ctx->nIF) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (boundary_faces && ctx->nBnd)
        CU(cudaMemcpyAsync(boundary_faces, src + ctx->nIF, size_t(ctx->nBnd) * sizeof(double), cudaMemcpyDeviceToHost,
                           ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SSAM_OK;
}

ssam_status ssam_face_field_upload(ssam_ctx* ctx, int ff, const double* internal_faces, const double* boundary_faces) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx && ctx->haveMesh, SSAM_ERR_INVALID, "mesh not set");
    CHECK(ff >= 0 && ff < SSAM_FF_COUNT, SSAM_ERR_INVALID, "bad face-field id");
    TRY(ensureFaceField(ctx, ff));
    if (internal_faces && ctx->nIF) {
        if (ctx->faceKey) {  // tile-compact layout: device face g holds the caller's face faceKey[g]
            TRY(ensureStage(ctx, size_t(ctx->nIF)));
            CU(cudaMemcpyAsync(ctx->stage, internal_faces, size_t(ctx->nIF) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            LAUNCH(k_gather_by_key, gridFor(ctx->nIF, 256), 256, ctx->stage, ctx->faceKey, ctx->nIF, ctx->faceField[ff]);
            CU(cudaStreamSynchronize(ctx->stream));
        } else {
            CU(cudaMemcpyAsync(ctx->faceField[ff], internal_faces, size_t(ctx->nIF) * sizeof(double), cudaMemcpyHostToDevice,
                               ctx->stream));
        }
    }
    if (boundary_faces && ctx->nBnd)
        CU(cudaMemcpyAsync(ctx->faceField[ff] + ctx->nIF, boundary_faces, size_t(ctx->nBnd) * sizeof(double),
                           cudaMemcpyHostToDevice, ctx->stream));
    ctx->faceFieldSet[ff] = true;
    return finish(ctx);
}

ssam_status ssam_alpha_compression_coeff(ssam_ctx* ctx, double cAlpha, double icAlpha, double scAlpha) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx && ctx->haveMesh, SSAM_ERR_INVALID, "mesh not set");
    CHECK(ctx->faceFieldSet[SSAM_FF_PHI], SSAM_ERR_MISSING_FIELD, "field not available: phi (ssam_face_field_upload)");
    if (icAlpha > 0) TRY(need(ctx, SSAM_F_U, "U"));
    if (scAlpha > 0) TRY(need(ctx, SSAM_F_GRADU, "gradU (call ssam_grad_u first)"));
    TRY(ensureFaceField(ctx, SSAM_FF_PHIC));
    if (scAlpha > 0 && ctx->nProcPatches > 0 && !ctx->procNbrCSet) {
        // processorPolyPatch::neighbFaceCellCentres(), once per mesh
        PlanePtrs snd, rcv;
        for (int k = 0; k < 16; ++k) snd.p[k] = rcv.p[k] = nullptr;
        for (int k = 0; k < 3; ++k) {
            if (!ctx->procNbrC[k]) {
                TRY(devAlloc(ctx, &ctx->procNbrC[k], size_t(ctx->nBnd)));
                CU(cudaMemsetAsync(ctx->procNbrC[k], 0, size_t(ctx->nBnd) * sizeof(double), ctx->stream));
            }
            snd.p[k] = ctx->Cpl[k];
            rcv.p[k] = ctx->procNbrC[k] - ctx->nCells;
        }
        TRY(haloExchangePlanes(ctx, snd, rcv, 3, ctx->stream));
        ctx->procNbrCSet = true;
    }
    PhicArgs a;
    std::memset(&a, 0, sizeof(a));
    a.nCells = ctx->nCells;
    a.nIF = ctx->nIF;
    a.nBnd = ctx->nBnd;
    a.geo = ctx->geo;
    a.owner = ctx->owner;
    a.neighbour = ctx->neighbour;
    a.bndFlags = ctx->bndFlags;
    for (int k = 0; k < 3; ++k) {
        a.C[k] = ctx->Cpl[k];
        a.nbrC[k] = ctx->procNbrC[k];
        a.U[k] = ctx->plane[SSAM_F_U][k];
    }
    for (int k = 0; k < 9; ++k) a.g[k] = ctx->plane[SSAM_F_GRADU][k];
    a.phi = ctx->faceField[SSAM_FF_PHI];
    a.phic = ctx->faceField[SSAM_FF_PHIC];
    a.cAlpha = cAlpha;
    a.icAlpha = icAlpha;
    a.scAlpha = scAlpha;
    LAUNCH(k_phic, gridFor(ctx->nFaces, 256), 256, a);
    ctx->faceFieldSet[SSAM_FF_PHIC] = true;
    return finish(ctx);
}

double* ssam_face_field_device_ptr(ssam_ctx* ctx, int ff) {
    return (ctx && ff >= 0 && ff < SSAM_FF_COUNT) ? ctx->faceField[ff] : nullptr;
}

static ssam_status mixtureFaceVisc(ssam_ctx* ctx, const ssam_mixture_params* p, bool nuf) {
    CHECK(ctx && p, SSAM_ERR_INVALID, "null argument");
    TRY(need(ctx, SSAM_F_ALPHA1, "alpha1"));
    TRY(need(ctx, SSAM_F_NU1, "nu1 (viscosityModel::correct() of phase 1)"));
    TRY(need(ctx, SSAM_F_NU2, "nu2 (viscosityModel::correct() of phase 2)"));
    const int ff = nuf ? SSAM_FF_NUF : SSAM_FF_MUF;
    TRY(ensureFaceField(ctx, ff));
    FaceViscArgs a;
    a.nCells = ctx->nCells;
    a.nIF = ctx->nIF;
    a.nBnd = ctx->nBnd;
    a.geo = ctx->geo;
    a.owner = ctx->owner;
    a.neighbour = ctx->neighbour;
    a.bndFlags = ctx->bndFlags;
    a.alpha = ctx->plane[SSAM_F_ALPHA1][0];
    a.nu1 = ctx->plane[SSAM_F_NU1][0];
    a.nu2 = ctx->plane[SSAM_F_NU2][0];
    a.rho1 = p->rho1;
    a.rho2 = p->rho2;
    a.out = ctx->faceField[ff];
    if (nuf) {
        LAUNCH((k_face_visc<true>), gridFor(ctx->nFaces, 256), 256, a);
    } else {
        LAUNCH((k_face_visc<false>), gridFor(ctx->nFaces, 256), 256, a);
    }
    ctx->faceFieldSet[ff] = true;
    return finish(ctx);
}
ssam_status ssam_face_interpolate(ssam_ctx* ctx, int field, int ff) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx, SSAM_ERR_INVALID, "null context");
    CHECK(field >= 0 && field < SSAM_F_COUNT && ncompOf(field) == 1, SSAM_ERR_INVALID, "scalar field id expected");
    CHECK(ff >= 0 && ff < SSAM_FF_COUNT, SSAM_ERR_INVALID, "bad face field id");
    TRY(need(ctx, field, "field to interpolate"));
    TRY(ensureFaceField(ctx, ff));
    FaceViscArgs a;
    std::memset(&a, 0, sizeof(a));
    a.nCells = ctx->nCells;
    a.nIF = ctx->nIF;
    a.nBnd = ctx->nBnd;
    a.geo = ctx->geo;
    a.owner = ctx->owner;
    a.neighbour = ctx->neighbour;
    a.bndFlags = ctx->bndFlags;
    a.alpha = ctx->plane[field][0];
    a.out = ctx->faceField[ff];
    LAUNCH(k_face_interpolate, gridFor(ctx->nFaces, 256), 256, a);
    ctx->faceFieldSet[ff] = true;
    return finish(ctx);
}
ssam_status ssam_mixture_muf(ssam_ctx* ctx, const ssam_mixture_params* p) { return mixtureFaceVisc(ctx, p, false); }
ssam_status ssam_mixture_nuf(ssam_ctx* ctx, const ssam_mixture_params* p) { return mixtureFaceVisc(ctx, p, true); }

ssam_status ssam_field_min_max(ssam_ctx* ctx, int field, double* min_value, double* max_value) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx && min_value && max_value, SSAM_ERR_INVALID, "null argument");
    CHECK(field >= 0 && field < SSAM_F_COUNT && ncompOf(field) == 1, SSAM_ERR_INVALID, "scalar field id expected");
    TRY(need(ctx, field, "min/max of a field that was never set"));
    const int nBlocks = std::min(gridFor(ctx->nTot, 256), 148 * 8);
    if (!ctx->minmaxScratch) TRY(devAlloc(ctx, &ctx->minmaxScratch, size_t(2 * 148 * 8 + 2)));
    double* out = ctx->minmaxScratch + 2 * 148 * 8;
    LAUNCH(k_minmax, nBlocks, 256, ctx->plane[field][0], ctx->nTot, ctx->minmaxScratch);
    LAUNCH(k_minmax_final, 1, 256, ctx->minmaxScratch, nBlocks, out);
    if (!ctx->localPeers.empty() && ctx->nRanks > 1) {
        return fail(ctx, SSAM_ERR_COMM, "ssam_field_min_max across an in-process rank group: reduce the per-rank values "
                                        "on the host");
    } else if (ctx->comm && ctx->nRanks > 1) {
        NcclApi& N = nccl();
        NC(N.GroupStart());
        NC(N.AllReduce(out, out, 1, ncclFloat64, ncclMin, ctx->comm, ctx->stream));
        NC(N.AllReduce(out + 1, out + 1, 1, ncclFloat64, ncclMax, ctx->comm, ctx->stream));
        NC(N.GroupEnd());
        ctx->launches++;
    }
    double h[2];
    CU(cudaMemcpyAsync(h, out, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    *min_value = h[0];
    *max_value = h[1];
    return SSAM_OK;
}

ssam_status ssam_interface_nhatf_download(ssam_ctx* ctx, double* internal_faces, double* boundary_faces) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    return ssam_face_field_download(ctx, SSAM_FF_NHATF, internal_faces, boundary_faces);
}
double* ssam_interface_nhatf_device_ptr(ssam_ctx* ctx) { return ssam_face_field_device_ptr(ctx, SSAM_FF_NHATF); }

ssam_status ssam_interface_delta_n(ssam_ctx* ctx, double* deltaN) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx && deltaN, SSAM_ERR_INVALID, "null argument");
    CHECK(ctx->haveMesh, SSAM_ERR_INVALID, "mesh not set");
    double sn[2] = {ctx->sumV, double(ctx->nCells)};
    if (ctx->comm && ctx->nRanks > 1) {
        // gAverage: sum and count over all ranks
        double* tmp = nullptr;
        TRY(devAlloc(ctx, &tmp, 2));
        CU(cudaMemcpyAsync(tmp, sn, sizeof(sn), cudaMemcpyHostToDevice, ctx->stream));
        NC(nccl().AllReduce(tmp, tmp, 2, ncclFloat64, ncclSum, ctx->comm, ctx->stream));
        CU(cudaMemcpyAsync(sn, tmp, sizeof(sn), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        cudaFree(tmp);
    }
    *deltaN = 1e-8 / std::pow(sn[0] / sn[1], 1.0 / 3.0);
    return SSAM_OK;
}

ssam_status ssam_friction_correct(ssam_ctx* ctx, const ssam_stick_params* p) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx && p, SSAM_ERR_INVALID, "null argument");
    CHECK(p->model == SSAM_STICK_CONSTANT || p->model == SSAM_STICK_EXPONENTIAL, SSAM_ERR_UNKNOWN_MODEL,
          "Unknown stickModel type");
    TRY(need(ctx, SSAM_F_MUEFF, "muEff"));
    TRY(need(ctx, SSAM_F_EPSDOT, "epsilonDotEq"));
    TRY(ensureField(ctx, SSAM_F_QVISC));
    QviscStage s{p->betaTQ, ctx->plane[SSAM_F_MUEFF][0], ctx->plane[SSAM_F_EPSDOT][0], ctx->plane[SSAM_F_QVISC][0]};
    LAUNCH((k_map<QviscStage>), mapGrid(ctx), 256, ctx->nTot, s);
    ctx->isSet[SSAM_F_QVISC] = true;
    if (p->fricPatch >= 0) TRY(launchQfric(ctx, *p));
    return finish(ctx);
}

ssam_status ssam_friction_patch_centre(ssam_ctx* ctx, double centre[3], double* area) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx && centre && area, SSAM_ERR_INVALID, "null argument");
    CHECK(ctx->fricPatchCached >= 0, SSAM_ERR_INVALID, "no friction patch evaluated yet");
    double s[4];
    CU(cudaMemcpyAsync(s, ctx->fricSums + 4, sizeof(s), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    centre[0] = s[0] / s[3];
    centre[1] = s[1] / s[3];
    centre[2] = s[2] / s[3];
    *area = s[3];
    return SSAM_OK;
}

// the high-priority side stream of the multi-rank schedules and the events that tie it to the main stream
static ssam_status ensureSideStream(ssam_ctx* ctx) {
    if (ctx->commStream) return SSAM_OK;
    // highest priority: the side stream's small kernels must get SM slots as soon as CTAs of the
    // interior-tile kernel retire, not after its whole grid has been dispatched
    int prLo = 0, prHi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&prLo, &prHi));
    CU(cudaStreamCreateWithPriority(&ctx->commStream, cudaStreamNonBlocking, prHi));
    CU(cudaEventCreateWithFlags(&ctx->evMainToComm, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ctx->evCommToMain, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ctx->evHaloIn, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ctx->evCoupledDone, cudaEventDisableTiming));
    return SSAM_OK;
}

// ---- fused update ---------------------------------------------------------------------------------------------
// Host-staged pipeline (ssam_update_fused_host): the cell kernel runs chunk by chunk (contiguous tile
// ranges); chunk k starts when the inputs of chunks <= k+1 have landed and signals evOut[k] when done.
struct ChunkPlan {
    int nChunks = 0, tilesPerChunk = 0;
    std::vector<cudaEvent_t> evIn, evOut;
    std::vector<int> waitIn;  // compute chunk k may start after upload chunk waitIn[k] (empty: k+1)
    const int* list = nullptr;  // launch order of the tiles (device array of nTiles entries); nullptr = identity
};

static ssam_status updateFusedAsync(ssam_ctx* ctx, const ssam_update_params* p, int flags,
                                    const ChunkPlan* plan = nullptr) {
    CHECK(ctx && p, SSAM_ERR_INVALID, "null argument");
    TRY(checkViscModel(ctx, &p->visc1));
    TRY(checkViscModel(ctx, &p->visc2));
    CHECK(p->visc2.model == SSAM_VISC_NEWTONIAN, SSAM_ERR_INVALID,
          "fused path: phase 2 must be Newtonian (two viscoplastic phases clash on sigmaf/sigmay, SURVEY D-2); "
          "use the staged calls");
    TRY(need(ctx, SSAM_F_U, "U"));
    TRY(need(ctx, SSAM_F_T, "temp"));
    TRY(need(ctx, SSAM_F_ALPHA1, "alpha1"));
    UpdateArgs a;
    fillMeshArgs(ctx, a);
    TRY(unitsShift(ctx, p->mix.k1, &a.kshift1));
    TRY(unitsShift(ctx, p->mix.k2, &a.kshift2));
    const int model1 = p->visc1.model;
    const bool hasSigma = model1 == SSAM_VISC_SHEPPARD_WRIGHT || model1 == SSAM_VISC_FKS;
    const int outs[] = {SSAM_F_EPSDOT, SSAM_F_NU1, SSAM_F_NU2, SSAM_F_NU,  SSAM_F_MUEFF,
                        SSAM_F_RHO,    SSAM_F_CP,  SSAM_F_K,   SSAM_F_QVISC};
    for (int f : outs) TRY(ensureField(ctx, f));
    if (hasSigma) {
        TRY(ensureField(ctx, SSAM_F_SIGMAF));
        TRY(ensureField(ctx, SSAM_F_SIGMAY));
    }
    if (model1 == SSAM_VISC_NORTON_HOFF) TRY(ensureField(ctx, SSAM_F_STRAINRATE));
    if (p->stick.fricPatch >= 0) TRY(ensureField(ctx, SSAM_F_QFRIC));
    if (!ctx->fricSums) TRY(devAlloc(ctx, &ctx->fricSums, 8));
    if (flags & 1) {
        TRY(ensureField(ctx, SSAM_F_GRADU));
        for (int k = 0; k < 9; ++k) a.gradU[k] = ctx->plane[SSAM_F_GRADU][k];
        a.writeGrad = 1;
    }
    a.T = ctx->plane[SSAM_F_T][0];
    a.alpha = ctx->plane[SSAM_F_ALPHA1][0];
    a.eps = ctx->plane[SSAM_F_EPSDOT][0];
    a.sr = ctx->plane[SSAM_F_STRAINRATE][0];
    a.nu1 = ctx->plane[SSAM_F_NU1][0];
    a.nu2 = ctx->plane[SSAM_F_NU2][0];
    a.sigmaf = ctx->plane[SSAM_F_SIGMAF][0];
    a.sigmay = ctx->plane[SSAM_F_SIGMAY][0];
    a.nu = ctx->plane[SSAM_F_NU][0];
    a.mu = ctx->plane[SSAM_F_MUEFF][0];
    a.rho = ctx->plane[SSAM_F_RHO][0];
    a.cp = ctx->plane[SSAM_F_CP][0];
    a.k = ctx->plane[SSAM_F_K][0];
    a.qvisc = ctx->plane[SSAM_F_QVISC][0];
    a.prm = *p;
    a.d1 = deriveVisc(p->visc1);
    a.pfDist = ctx->pfDist;
    if (const char* pf = std::getenv("SSAM_PF")) a.pfDist = std::atoi(pf);
    a.pfDistU = a.pfDist + ctx->pfBandwidthTiles;

    // nu2 (Newtonian, constant): written only when the value changes or something else wrote the plane
    if (!ctx->nu2Filled || ctx->nu2FilledWith != p->visc2.nu) {
        FillStage fs{p->visc2.nu, ctx->plane[SSAM_F_NU2][0]};
        LAUNCH((k_map<FillStage>), mapGrid(ctx), 256, ctx->nTot, fs);
        ctx->nu2Filled = true;
        ctx->nu2FilledWith = p->visc2.nu;
    }
    // Multi-GPU schedules (tiled kernels).  The tiles are split into "interior" ones and "coupled" ones (a
    // cell of the tile has a processor face, so its gradient needs the neighbour rank's U).
    //  (a) peer memory (default), everything on ONE stream:
    //        side stream: push(U,temp,alpha1) -> wait flags + unpack -> coupled tiles -> push(epsilonDotEq[,..])
    //        -> wait flags + unpack; main stream: all interior tiles, then (after the side stream) the boundary faces
    //      push = one small kernel storing straight into the neighbours' IPC-mapped receive slots over NVLink and
    //      publishing a flag; wait = a stream memory operation.  The neighbours' values travel while the interior
    //      tiles (75-99 % of the step) run, no library kernel shares the SMs with them, and a rank blocks only when
    //      a neighbour is really behind by more than that.
    //  (b) NCCL (SSAM_HALO=nccl or a neighbour that cannot be mapped): round-1 schedule, exchanges and coupled tiles
    //      on a high-priority second stream beside the interior tiles.
    const bool haloInputs = ctx->nProcPatches && (flags & 2);
    const bool split = ctx->nProcPatches > 0 && ctx->cellKernel >= 1;
    CHECK(ctx->nProcPatches == 0 || ctx->localPeers.empty() || !ctx->blocking, SSAM_ERR_INVALID,
          "in-process rank group: collective calls need ssam_set_blocking(ctx, 0) on every rank (enqueue all ranks, "
          "then synchronise)");
    CHECK(ctx->nProcPatches == 0 || ctx->localPeers.empty() || ctx->fusedPrepared, SSAM_ERR_INVALID,
          "in-process rank group: call ssam_prepare_fused on every rank before the first collective update");
    if (ctx->nProcPatches) TRY(ensureHaloTransport(ctx));
    const bool peer = split && ctx->p2p.state == 1;
    if (split) TRY(ensureSideStream(ctx));
    int outs2[3], n2 = 0;
    outs2[n2++] = SSAM_F_EPSDOT;
    if (model1 == SSAM_VISC_NORTON_HOFF) outs2[n2++] = SSAM_F_STRAINRATE;
    if (flags & 1) outs2[n2++] = SSAM_F_GRADU;
    const int ins[3] = {SSAM_F_U, SSAM_F_T, SSAM_F_ALPHA1};
    ctx->haloBytesLastUpdate = 0;
    // profiling: CUDA events around every tile-kernel launch of this update (their sum = the cell kernel's time)
    auto profBegin = [&](cudaEvent_t& e0) -> ssam_status {
        e0 = nullptr;
        if (!ctx->profile) return SSAM_OK;
        CU(cudaEventCreate(&e0));
        CU(cudaEventRecord(e0, ctx->stream));
        return SSAM_OK;
    };
    auto profEnd = [&](cudaEvent_t e0) -> ssam_status {
        if (!e0) return SSAM_OK;
        cudaEvent_t e1 = nullptr;
        CU(cudaEventCreate(&e1));
        CU(cudaEventRecord(e1, ctx->stream));
        ctx->profEvents.emplace_back(e0, e1);
        return SSAM_OK;
    };
    auto launchInterior = [&](int first, int count) -> ssam_status {  // interior tiles [first, first + count)
        if (count <= 0) return SSAM_OK;
        if (ctx->interiorRange0 >= 0) {
            UpdateArgs ar = a;
            ar.tileOffset = ctx->interiorRange0 + first;
            return launchCellsTiled(ctx, ar, model1, nullptr, count, ctx->stream);
        }
        return launchCellsTiled(ctx, a, model1, ctx->tileListInterior + first, count, ctx->stream);
    };
    cudaEvent_t pe = nullptr;
    if (peer) {
        PlanePtrs plIn, plOut;
        int ncIn = 0, ncOut = 0;
        TRY(planesOf(ctx, ins, 3, plIn, ncIn));
        TRY(planesOf(ctx, outs2, n2, plOut, ncOut));
        unsigned long long seqIn = 0, seqOut = 0;
        // Everything that depends on another rank runs on the high-priority side stream beside ONE launch of all
        // interior tiles: push, stream-ordered wait and unpack of the inputs, the coupled tiles, push / wait / unpack of
        // the outputs.  None of these pieces ever spins (an NCCL send/recv kernel does), the small kernels get SM slots
        // as interior CTAs retire, and the main stream has no launch boundary inside the step:
        //   main: interior tiles ............................................................ | boundary faces
        //   side: push(in) wait unpack | coupled tiles | push(out) wait unpack ---------------^
        // (Measured at 2 GPUs on the 64M-cell box, profiles/r2_scaling.md: splitting the interior launch in two around
        // the coupled tiles on one stream costs ~60 us of launch-boundary bubbles per step.)
        cudaStream_t side = ctx->commStream;
        CU(cudaEventRecord(ctx->evMainToComm, ctx->stream));
        CU(cudaStreamWaitEvent(side, ctx->evMainToComm, 0));
        if (haloInputs) {
            TRY(haloPushPeer(ctx, plIn, ncIn, side, &seqIn));
            TRY(haloWaitUnpackPeer(ctx, plIn, ncIn, seqIn, side));
        }
        TRY(profBegin(pe));
        TRY(launchInterior(0, ctx->nTilesInterior));
        TRY(launchCellsTiled(ctx, a, model1, ctx->tileListCoupled, ctx->nTilesCoupled, side));
        TRY(haloPushPeer(ctx, plOut, ncOut, side, &seqOut));
        TRY(haloWaitUnpackPeer(ctx, plOut, ncOut, seqOut, side));
        CU(cudaEventRecord(ctx->evCommToMain, side));
        CU(cudaStreamWaitEvent(ctx->stream, ctx->evCommToMain, 0));
        TRY(profEnd(pe));
    } else if (split) {
        TRY(profBegin(pe));
        CU(cudaEventRecord(ctx->evMainToComm, ctx->stream));
        CU(cudaStreamWaitEvent(ctx->commStream, ctx->evMainToComm, 0));
        if (haloInputs) TRY(haloExchangeFields(ctx, ins, 3, ctx->commStream));
        TRY(launchInterior(0, ctx->nTilesInterior));
        TRY(launchCellsTiled(ctx, a, model1, ctx->tileListCoupled, ctx->nTilesCoupled, ctx->commStream));
        TRY(haloExchangeFields(ctx, outs2, n2, ctx->commStream));
        CU(cudaEventRecord(ctx->evCommToMain, ctx->commStream));
        CU(cudaStreamWaitEvent(ctx->stream, ctx->evCommToMain, 0));
        TRY(profEnd(pe));
    } else {
        if (haloInputs) TRY(haloExchangeFields(ctx, ins, 3, ctx->stream));
        TRY(profBegin(pe));
        if (plan) {
            for (int k = 0; k < plan->nChunks; ++k) {
                const int t0 = k * plan->tilesPerChunk, t1 = std::min(ctx->nTiles, t0 + plan->tilesPerChunk);
                const int dep = plan->waitIn.empty() ? std::min(k + 1, plan->nChunks - 1) : plan->waitIn[size_t(k)];
                CU(cudaStreamWaitEvent(ctx->stream, plan->evIn[size_t(dep)], 0));
                TRY(launchCellsTiled(ctx, a, model1, (plan->list ? plan->list : ctx->tileIota) + t0, t1 - t0, ctx->stream));
                CU(cudaEventRecord(plan->evOut[k], ctx->stream));
            }
        } else if (ctx->cellKernel >= 1) {
            TRY(launchCellsTiled(ctx, a, model1, ctx->tileOrder, ctx->nTiles, ctx->stream));
        } else {
            TRY((launchCellsBoundary<MODE_FUSED>(ctx, a, model1, false)));
        }
        TRY(profEnd(pe));
        if (ctx->nProcPatches) TRY(haloExchangeFields(ctx, outs2, n2, ctx->stream));
    }
    if (ctx->profile) ctx->profUpdates++;
    TRY((launchCellsBoundary<MODE_FUSED>(ctx, a, model1, true)));
    for (int f : outs) ctx->isSet[f] = true;
    if (hasSigma) ctx->isSet[SSAM_F_SIGMAF] = ctx->isSet[SSAM_F_SIGMAY] = true;
    if (model1 == SSAM_VISC_NORTON_HOFF) ctx->isSet[SSAM_F_STRAINRATE] = true;
    if (flags & 1) ctx->isSet[SSAM_F_GRADU] = true;
    if (p->stick.fricPatch >= 0) TRY(launchQfric(ctx, p->stick));
    return SSAM_OK;
}

// Everything an update allocates or sets up lazily, without enqueuing any exchange: fields, the friction scratch, the
// halo transport.  Optional for one-process-per-GPU runs; REQUIRED before the first collective call of an in-process
// rank group, where a host-blocking operation (allocation, pageable copy, lazy module load) issued while this rank's
// stream is parked on a wait for a rank that the same host thread has not enqueued yet would deadlock.
ssam_status ssam_prepare_fused(ssam_ctx* ctx, const ssam_update_params* p, int flags) {
    if (ctx) cudaSetDevice(ctx->device);
    CHECK(ctx && p && ctx->haveMesh, SSAM_ERR_INVALID, "mesh not set / null params");
    const int model1 = p->visc1.model;
    const int outs[] = {SSAM_F_EPSDOT, SSAM_F_NU1, SSAM_F_NU2, SSAM_F_NU,  SSAM_F_MUEFF, SSAM_F_RHO,
                        SSAM_F_CP,     SSAM_F_K,   SSAM_F_QVISC, SSAM_F_SIGMAF, SSAM_F_SIGMAY};
    for (int f : outs) TRY(ensureField(ctx, f));
    if (model1 == SSAM_VISC_NORTON_HOFF) TRY(ensureField(ctx, SSAM_F_STRAINRATE));
    if (flags & 1) TRY(ensureField(ctx, SSAM_F_GRADU));
    if (p->stick.fricPatch >= 0) TRY(ensureField(ctx, SSAM_F_QFRIC));
    if (!ctx->fricSums) TRY(devAlloc(ctx, &ctx->fricSums, 8));
    if (ctx->cellKernel == 2 && !ctx->pipeSched) {
        TRY(devAlloc(ctx, &ctx->pipeSched, 16));
        CU(cudaMemsetAsync(ctx->pipeSched, 0, 16 * sizeof(int), ctx->stream));
    }
    for (int k = 0; k < 3 && ctx->nProcPatches; ++k)
        if (!ctx->procNbrC[k]) {
            TRY(devAlloc(ctx, &ctx->procNbrC[k], size_t(ctx->nBnd)));
            CU(cudaMemsetAsync(ctx->procNbrC[k], 0, size_t(ctx->nBnd) * sizeof(double), ctx->stream));
        }
    TRY(ensureFaceField(ctx, SSAM_FF_PHI));
    TRY(ensureFaceField(ctx, SSAM_FF_PHIC));
    if (ctx->nProcPatches) {
        TRY(ensureHaloTransport(ctx));
        TRY(ensureSideStream(ctx));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->fusedPrepared = true;
    return SSAM_OK;
}

ssam_status ssam_update_fused(ssam_ctx* ctx, const ssam_update_params* p, int flags) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    TRY(updateFusedAsync(ctx, p, flags));
    return finish(ctx);
}

static ssam_status updateFusedHostSimple(ssam_ctx* ctx, const ssam_update_params* p, const ssam_host_io* io) {
    struct In {
        int f;
        const double *i, *b;
    } ins[] = {{SSAM_F_U, io->U_i, io->U_b},
               {SSAM_F_T, io->T_i, io->T_b},
               {SSAM_F_ALPHA1, io->alpha1_i, io->alpha1_b},
               {SSAM_F_P, io->p_i, io->p_b}};
    TRY(ensureStage(ctx, size_t(ctx->nTot) * 3));
    for (const In& in : ins) {
        if (!in.i) continue;
        TRY(ensureField(ctx, in.f));
        const int nc = ncompOf(in.f);
        TRY(uploadAoS(ctx, in.i, ctx->nCells, nc, ctx->plane[in.f], 0));
        if (in.b) TRY(uploadAoS(ctx, in.b, ctx->nBnd, nc, ctx->plane[in.f], ctx->nCells));
        ctx->isSet[in.f] = true;
    }
    TRY(updateFusedAsync(ctx, p, ctx->nProcPatches ? 2 : 0));
    for (int o = 0; o < io->nOut; ++o) {
        const int f = io->outFields[o];
        TRY(need(ctx, f, "requested output was not produced by this update"));
        if (io->out_i && io->out_i[o]) TRY(downloadAoS(ctx, io->out_i[o], ctx->nCells, 1, ctx->plane[f], 0));
        if (io->out_b && io->out_b[o] && ctx->nBnd)
            CU(cudaMemcpyAsync(io->out_b[o], ctx->plane[f][0] + ctx->nCells, size_t(ctx->nBnd) * sizeof(double),
                               cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return SSAM_OK;
}

// Pipelined variant (single GPU, tiled kernel): PCIe is full duplex, so the upload of chunk k+2's
// inputs, the cell kernel on chunk k+1 and the download of chunk k's outputs run concurrently on three
// streams.  The gather of a chunk reaches at most `bandwidthTiles` tiles ahead, hence the k+1 dependency.
static ssam_status updateFusedHostPipelined(ssam_ctx* ctx, const ssam_update_params* p, const ssam_host_io* io,
                                            int nChunks) {
    if (!ctx->h2dStream) {
        CU(cudaStreamCreateWithFlags(&ctx->h2dStream, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&ctx->d2hStream, cudaStreamNonBlocking));
    }
    ChunkPlan plan;
    plan.nChunks = nChunks;
    plan.tilesPerChunk = (ctx->nTiles + nChunks - 1) / nChunks;
    plan.evIn.resize(nChunks);
    plan.evOut.resize(nChunks);
    for (int k = 0; k < nChunks; ++k) {
        CU(cudaEventCreateWithFlags(&plan.evIn[k], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&plan.evOut[k], cudaEventDisableTiming));
    }
    cudaEvent_t evStart, evTail;
    CU(cudaEventCreateWithFlags(&evStart, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&evTail, cudaEventDisableTiming));
    const int fIn[4] = {SSAM_F_U, SSAM_F_T, SSAM_F_ALPHA1, SSAM_F_P};
    const double* hi[4] = {io->U_i, io->T_i, io->alpha1_i, io->p_i};
    const double* hb[4] = {io->U_b, io->T_b, io->alpha1_b, io->p_b};
    for (int i = 0; i < 4; ++i)
        if (hi[i]) TRY(ensureField(ctx, fIn[i]));
    const size_t chunkCells = size_t(plan.tilesPerChunk) * TILE;
    const bool permuted = ctx->perm != nullptr;
    std::vector<int> outWait;  // download chunk j may start after compute chunk outWait[j]
    int nScalarIn = 0;
    for (int i = 1; i < 4; ++i) nScalarIn += hi[i] ? 1 : 0;
    if (permuted) {
        // tile-compact layout: the caller's chunk [c0, c1) is scattered over the device order.  Compute chunk k needs
        // every caller cell its tiles (and their neighbours) hold; download chunk j needs every tile that holds
        // one of its cells.  Both are monotone enough for a pipeline because runs of 64 caller cells stay together
        // and clusters only reach a few mesh layers.
        if (ctx->tileMaxCallerCell.empty()) {
            int* d = nullptr;
            TRY(devAlloc(ctx, &d, size_t(ctx->nTiles)));
            k_tile_max_caller_cell<<<ctx->nTiles, 128, 0, ctx->stream>>>(ctx->ownerStart, ctx->neighbour, ctx->extraStart,
                                                                       ctx->extra, ctx->perm, ctx->nCells, ctx->nIF,
                                                                       ctx->nTiles, d);
            ctx->launches++;
            ctx->tileMaxCallerCell.resize(size_t(ctx->nTiles));
            CU(cudaMemcpyAsync(ctx->tileMaxCallerCell.data(), d, size_t(ctx->nTiles) * sizeof(int), cudaMemcpyDeviceToHost,
                               ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            cudaFree(d);
        }
        if (ctx->pipeListHost.empty()) {
            // tiles by their first (smallest) caller cell
            const int nT = ctx->nTiles;
            std::vector<int> first(size_t(nT), ctx->nCells);
            for (int i = 0; i < ctx->nCells; ++i) first[size_t(i / TILE)] = std::min(first[size_t(i / TILE)], ctx->permHost[size_t(i)]);
            ctx->pipeListHost.resize(size_t(nT));
            for (int t = 0; t < nT; ++t) ctx->pipeListHost[size_t(t)] = t;
            std::stable_sort(ctx->pipeListHost.begin(), ctx->pipeListHost.end(),
                             [&](int x, int y) { return first[size_t(x)] < first[size_t(y)]; });
            ctx->pipePosOfTile.resize(size_t(nT));
            for (int q = 0; q < nT; ++q) ctx->pipePosOfTile[size_t(ctx->pipeListHost[size_t(q)])] = q;
            TRY(devAlloc(ctx, &ctx->pipeList, size_t(nT)));
            CU(cudaMemcpyAsync(ctx->pipeList, ctx->pipeListHost.data(), size_t(nT) * sizeof(int), cudaMemcpyHostToDevice,
                               ctx->stream));
            TRY(devAlloc(ctx, &ctx->descPipe, size_t(nT)));
            LAUNCH(k_tile_desc_list, gridFor(nT, 256), 256, ctx->tileDesc, ctx->pipeList, nT, ctx->descPipe);
            CU(cudaStreamSynchronize(ctx->stream));
        }
        if (ctx->pipeChunks != nChunks) {
            ctx->pipeWaitIn.assign(size_t(nChunks), nChunks - 1);
            for (int k = 0; k < nChunks; ++k) {
                int m = 0;
                for (int q = k * plan.tilesPerChunk; q < std::min(ctx->nTiles, (k + 1) * plan.tilesPerChunk); ++q)
                    m = std::max(m, ctx->tileMaxCallerCell[size_t(ctx->pipeListHost[size_t(q)])]);
                ctx->pipeWaitIn[size_t(k)] = std::min(nChunks - 1, int(size_t(m) / chunkCells));
            }
            ctx->pipeOutWait.assign(size_t(nChunks), 0);
            for (int i = 0; i < ctx->nCells; ++i) {
                const size_t j = size_t(ctx->permHost[size_t(i)]) / chunkCells;
                ctx->pipeOutWait[j] = std::max(ctx->pipeOutWait[j], ctx->pipePosOfTile[size_t(i / TILE)] / plan.tilesPerChunk);
            }
            ctx->pipeChunks = nChunks;
        }
        plan.list = ctx->pipeList;
        plan.waitIn = ctx->pipeWaitIn;
        outWait = ctx->pipeOutWait;
    }
    // staging: uploads 2 x (3 + scalars) x chunk, downloads 2 x nOut x chunk (permuted layout only)
    const size_t upStage = 2 * chunkCells * size_t(3 + (permuted ? nScalarIn : 0));
    const size_t dnStage = permuted ? 2 * chunkCells * size_t(std::max(io->nOut, 1)) : 0;
    TRY(ensureStage(ctx, std::max(size_t(ctx->nBnd) * 3, upStage + dnStage)));
    double* const dnBase = ctx->stage + upStage;
    // order the copy streams after whatever the caller queued on the compute stream
    CU(cudaEventRecord(evStart, ctx->stream));
    CU(cudaStreamWaitEvent(ctx->h2dStream, evStart, 0));
    CU(cudaStreamWaitEvent(ctx->d2hStream, evStart, 0));
    cudaStream_t keep = ctx->stream;
    // boundary parts first (small), on the upload stream
    ctx->stream = ctx->h2dStream;
    ssam_status st = SSAM_OK;
    for (int i = 0; i < 4 && st == SSAM_OK; ++i)
        if (hi[i] && hb[i] && ctx->nBnd)
            st = uploadAoS(ctx, hb[i], ctx->nBnd, ncompOf(fIn[i]), ctx->plane[fIn[i]], ctx->nCells);
    PlanePtrs plU;
    for (int k = 0; k < 16; ++k) plU.p[k] = k < 3 ? ctx->plane[SSAM_F_U][k] : nullptr;
    for (int k = 0; k < nChunks && st == SSAM_OK; ++k) {
        const int c0 = int(std::min(size_t(ctx->nCells), size_t(k) * chunkCells));
        const int c1 = int(std::min(size_t(ctx->nCells), size_t(k + 1) * chunkCells));
        const int n = c1 - c0;
        if (n > 0) {
            double* stg = ctx->stage + size_t(k & 1) * (upStage / 2);
            cudaMemcpyAsync(stg, io->U_i + size_t(c0) * 3, size_t(n) * 3 * sizeof(double), cudaMemcpyHostToDevice,
                            ctx->h2dStream);
            if (permuted) {
                k_chunk_scatter<<<gridFor((long long)n * 3, 256), 256, 0, ctx->h2dStream>>>(stg, n, 3, plU, c0, ctx->invPerm);
                ctx->launches++;
                int slot = 0;
                for (int i = 1; i < 4; ++i) {
                    if (!hi[i]) continue;
                    double* ss = stg + chunkCells * size_t(3 + slot++);
                    cudaMemcpyAsync(ss, hi[i] + c0, size_t(n) * sizeof(double), cudaMemcpyHostToDevice, ctx->h2dStream);
                    PlanePtrs p1;
                    for (int q = 0; q < 16; ++q) p1.p[q] = q == 0 ? ctx->plane[fIn[i]][0] : nullptr;
                    k_chunk_scatter<<<gridFor(n, 256), 256, 0, ctx->h2dStream>>>(ss, n, 1, p1, c0, ctx->invPerm);
                    ctx->launches++;
                }
            } else {
                k_aos_to_soa_n<<<gridFor((long long)n * 3, 256), 256, 0, ctx->h2dStream>>>(stg, n, 3, plU, c0, nullptr);
                ctx->launches++;
                for (int i = 1; i < 4; ++i)
                    if (hi[i])
                        cudaMemcpyAsync(ctx->plane[fIn[i]][0] + c0, hi[i] + c0, size_t(n) * sizeof(double),
                                        cudaMemcpyHostToDevice, ctx->h2dStream);
            }
        }
        if (cudaEventRecord(plan.evIn[k], ctx->h2dStream) != cudaSuccess) st = SSAM_ERR_CUDA;
    }
    ctx->stream = keep;
    if (st != SSAM_OK) return fail(ctx, st, "host-staged pipeline: upload failed");
    for (int i = 0; i < 4; ++i)
        if (hi[i]) ctx->isSet[fIn[i]] = true;
    // the compute stream must not start before the boundary values are there: evIn[0] is recorded after them
    CU(cudaStreamWaitEvent(ctx->stream, plan.evIn[0], 0));
    TRY(updateFusedAsync(ctx, p, 0, &plan));
    CU(cudaEventRecord(evTail, ctx->stream));
    // downloads: internal parts chunk by chunk as the cell kernel finishes them (qfric and the boundary
    // parts are final only after the boundary / friction kernels, i.e. after evTail)
    for (int k = 0; k < nChunks; ++k) {
        const int c0 = int(std::min(size_t(ctx->nCells), size_t(k) * chunkCells));
        const int c1 = int(std::min(size_t(ctx->nCells), size_t(k + 1) * chunkCells));
        if (c1 <= c0) continue;
        CU(cudaStreamWaitEvent(ctx->d2hStream, plan.evOut[size_t(permuted ? outWait[size_t(k)] : k)], 0));
        for (int o = 0; o < io->nOut; ++o) {
            const int f = io->outFields[o];
            if (f == SSAM_F_QFRIC || !(io->out_i && io->out_i[o])) continue;
            TRY(need(ctx, f, "requested output was not produced by this update"));
            const double* src = ctx->plane[f][0] + c0;
            if (permuted) {
                double* g = dnBase + (size_t(k & 1) * size_t(io->nOut) + size_t(o)) * chunkCells;
                k_chunk_gather<<<gridFor(c1 - c0, 256), 256, 0, ctx->d2hStream>>>(g, c1 - c0, ctx->plane[f][0], c0,
                                                                                 ctx->invPerm);
                ctx->launches++;
                src = g;
            }
            CU(cudaMemcpyAsync(io->out_i[o] + c0, src, size_t(c1 - c0) * sizeof(double), cudaMemcpyDeviceToHost,
                               ctx->d2hStream));
        }
    }
    CU(cudaStreamWaitEvent(ctx->d2hStream, evTail, 0));
    for (int o = 0; o < io->nOut; ++o) {
        const int f = io->outFields[o];
        TRY(need(ctx, f, "requested output was not produced by this update"));
        if (f == SSAM_F_QFRIC && io->out_i && io->out_i[o]) {
            if (permuted) {
                cudaStream_t keep2 = ctx->stream;
                ctx->stream = ctx->d2hStream;  // downloadAoS applies the permutation (and synchronises its stream)
                const ssam_status s2 = downloadAoS(ctx, io->out_i[o], ctx->nCells, 1, ctx->plane[f], 0);
                ctx->stream = keep2;
                if (s2 != SSAM_OK) return s2;
            } else {
                CU(cudaMemcpyAsync(io->out_i[o], ctx->plane[f][0], size_t(ctx->nCells) * sizeof(double),
                                   cudaMemcpyDeviceToHost, ctx->d2hStream));
            }
        }
        if (io->out_b && io->out_b[o] && ctx->nBnd)
            CU(cudaMemcpyAsync(io->out_b[o], ctx->plane[f][0] + ctx->nCells, size_t(ctx->nBnd) * sizeof(double),
                               cudaMemcpyDeviceToHost, ctx->d2hStream));
    }
    CU(cudaStreamSynchronize(ctx->d2hStream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaStreamSynchronize(ctx->h2dStream));
    for (int k = 0; k < nChunks; ++k) {
        cudaEventDestroy(plan.evIn[k]);
        cudaEventDestroy(plan.evOut[k]);
    }
    cudaEventDestroy(evStart);
    cudaEventDestroy(evTail);
    return SSAM_OK;
}

ssam_status ssam_update_fused_host(ssam_ctx* ctx, const ssam_update_params* p, const ssam_host_io* io) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx && p && io, SSAM_ERR_INVALID, "null argument");
    CHECK(ctx->haveMesh, SSAM_ERR_INVALID, "mesh not set");
    CHECK(io->U_i && io->T_i && io->alpha1_i, SSAM_ERR_INVALID, "U, temp and alpha1 host buffers are required");
    for (int o = 0; o < io->nOut; ++o)
        CHECK(io->outFields[o] >= 0 && io->outFields[o] < SSAM_F_COUNT && ncompOf(io->outFields[o]) == 1,
              SSAM_ERR_INVALID, "host outputs must be scalar fields");
    // chunks of >= 1024 tiles, each at least as deep as the mesh bandwidth (so a chunk only gathers from
    // itself and its two neighbours)
    int nChunks = ctx->hostChunks >= 0 ? std::min(ctx->hostChunks, 64) : std::min(16, ctx->nTiles / 1024);
    // caller-order layout: the chunk dependency is "the next chunk"; the permuted layout derives it from the mesh
    if (!ctx->perm)
        while (nChunks > 1 && (ctx->nTiles + nChunks - 1) / nChunks < ctx->bandwidthTiles + 1) --nChunks;
    if (nChunks > 1 && ctx->nProcPatches == 0 && ctx->cellKernel >= 1)
        return updateFusedHostPipelined(ctx, p, io, nChunks);
    return updateFusedHostSimple(ctx, p, io);
}

// ---- device-math evaluation (tests) -----------------------------------------------------------------------------
ssam_status ssam_debug_math(ssam_ctx* ctx, int fn, const double* x, const double* y, double* out, int64_t n) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx && x && out && n >= 0, SSAM_ERR_INVALID, "null argument");
    if (n == 0) return SSAM_OK;
    double *dx = nullptr, *dy = nullptr, *dout = nullptr;
    TRY(devAlloc(ctx, &dx, size_t(n)));
    TRY(devAlloc(ctx, &dout, size_t(n)));
    CU(cudaMemcpyAsync(dx, x, size_t(n) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (y) {
        TRY(devAlloc(ctx, &dy, size_t(n)));
        CU(cudaMemcpyAsync(dy, y, size_t(n) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    LAUNCH(k_debug_math, gridFor(n, 256), 256, fn, dx, dy, dout, (long long)n);
    CU(cudaMemcpyAsync(out, dout, size_t(n) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(dx);
    cudaFree(dout);
    if (dy) cudaFree(dy);
    return SSAM_OK;
}

// ---- FP64 issue ceiling ------------------------------------------------------------------------------------------------
ssam_status ssam_measure_fp64_peak(ssam_ctx* ctx, double* fp64_instr_per_s) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx && fp64_instr_per_s, SSAM_ERR_INVALID, "null argument");
    CU(cudaSetDevice(ctx->device));
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
    const int grid = sms * 8, block = 256, iters = 1 << 15;
    double* out = nullptr;
    TRY(devAlloc(ctx, &out, size_t(grid) * block));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {  // the first repetition is the warm-up
        CU(cudaEventRecord(e0, ctx->stream));
        k_fp64_peak<<<grid, block, 0, ctx->stream>>>(iters, 1.0 + rep, out);
        ctx->launches++;
        CU(cudaEventRecord(e1, ctx->stream));
        CU(cudaEventSynchronize(e1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best = std::min(best, ms);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *fp64_instr_per_s = double(grid) * block * double(iters) * 8.0 / (double(best) * 1e-3);
    return SSAM_OK;
}

// ---- comm -----------------------------------------------------------------------------------------------------
ssam_status ssam_comm_unique_id(void* id_out) {
    ssam_ctx* ctx = nullptr;
    CHECK(id_out, SSAM_ERR_INVALID, "null id buffer");
    static_assert(sizeof(ncclUniqueId) <= SSAM_UNIQUE_ID_BYTES, "unique id size");
    if (!nccl().load()) return fail(nullptr, SSAM_ERR_COMM, nccl().error);
    ncclUniqueId id;
    NC(nccl().GetUniqueId(&id));
    std::memset(id_out, 0, SSAM_UNIQUE_ID_BYTES);
    std::memcpy(id_out, &id, sizeof(id));
    return SSAM_OK;
}

ssam_status ssam_comm_init(ssam_ctx* ctx, const void* id, int rank, int nRanks) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx && id, SSAM_ERR_INVALID, "null argument");
    CHECK(nRanks >= 1 && rank >= 0 && rank < nRanks, SSAM_ERR_INVALID, "bad rank / nRanks");
    if (!nccl().load()) return fail(ctx, SSAM_ERR_COMM, nccl().error);
    CU(cudaSetDevice(ctx->device));
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof(uid));
    NC(nccl().CommInitRank(&ctx->comm, nRanks, uid, rank));
    ctx->rank = rank;
    ctx->nRanks = nRanks;
    ctx->fricPatchCached = -1;
    return SSAM_OK;
}

ssam_status ssam_comm_init_local(ssam_ctx** ctxs, int nRanks) {
    ssam_ctx* ctx = nullptr;
    CHECK(ctxs && nRanks >= 1 && nRanks <= 32, SSAM_ERR_INVALID, "bad in-process rank group");
    for (int r = 0; r < nRanks; ++r) CHECK(ctxs[r] != nullptr, SSAM_ERR_INVALID, "null context in the rank group");
    for (int r = 0; r < nRanks; ++r) {
        ssam_ctx* c = ctxs[r];
        CHECK(c->comm == nullptr, SSAM_ERR_INVALID, "context already has an NCCL communicator");
        c->localPeers.assign(ctxs, ctxs + nRanks);
        c->rank = r;
        c->nRanks = nRanks;
        c->fricPatchCached = -1;
        p2pFree(c);
    }
    return SSAM_OK;
}

int64_t ssam_halo_bytes_last_update(const ssam_ctx* ctx) { return ctx ? int64_t(ctx->haloBytesLastUpdate) : 0; }

int ssam_halo_transport(const ssam_ctx* ctx) {
    if (!ctx || ctx->nProcPatches == 0 || ctx->p2p.state == 0) return 0;
    return ctx->p2p.state == 1 ? 1 : 2;
}

ssam_status ssam_halo_exchange(ssam_ctx* ctx, uint32_t mask) {
    if (ctx) cudaSetDevice(ctx->device);  // every entry point runs on the context's own GPU
    CHECK(ctx && ctx->haveMesh, SSAM_ERR_INVALID, "mesh not set");
    for (int f = 0; f < SSAM_F_COUNT; ++f)
        if (mask & (1u << f)) {
            TRY(need(ctx, f, "halo exchange of a field that is not set"));
            TRY(haloExchangeField(ctx, f));
        }
    return finish(ctx);
}

}  // extern "C"
