// dtcsim.cu -- sm_100a kernels + the C ABI declared in include/dtcsim.h.
//
// Kernel inventory (all complex128; no tensor cores on this path -- see DESIGN.md):
//   k_tile_stream<MODE>  fused  R_A|G -> D -> R_B|G  over 2^12-amplitude tiles, persistent and warp-specialised:
//                        TMA bulk / tensor-map loads and stores through a 3-stage shared-memory ring (mbarriers),
//                        table-builder warps, two compute warpgroups working in place (dtc_stream.cuh).  (hot kernel)
//   k_tile_pass<S2_LO>   the same pass with register-fed loads/stores, for tiles / bond patterns k_tile_stream does
//                        not take (general circuits).
//   k_frames             Philox4x32-10 Pauli-frame walk, one thread per trajectory.
//   k_probs / k_rdm / k_expect_z   warp-shuffle reductions for read-out; k_readout_small finishes the read-out of
//                        factorised circuits on a <= 3-qubit density matrix per trajectory (dtc_readout.cuh).
//   k_generic_*          one-thread-per-element fallbacks (n < 12, cross-checks).
//   k_dm_*               exact density-matrix primitives (small n).
//   k_shard_pack         gather / scatter for exchanges of a sharded state whose outgoing qubits are not the top bits.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdlib.h>

#include <new>
#include <string>
#include <vector>

#include "../../include/dtcsim.h"
#include "dtc_core.hpp"
#include "dtc_readout.cuh"
#include "dtc_dm.cuh"

// ------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int g_num_sms[16] = {0};
static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (expr);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            return fail(DTC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));       \
    } while (0)

// ---- per-device library state.  The library never leaves the caller's current device changed (DeviceGuard) and does
// not touch the process-wide default memory pool: program tables come from a PRIVATE stream-ordered pool (kept warm:
// cudaFree costs tens of milliseconds on a device whose address space holds multi-GiB state buffers, and run() creates
// and destroys one program per circuit) and are uploaded on a private non-blocking stream, so that creating the program
// of circuit i+1 never waits for the GPU work of circuit i that is queued on the caller's stream (a synchronous
// cudaMemcpy on the legacy default stream would).
#define DTC_MAX_DEVICES 16
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && device >= 0 && device != prev) err = cudaSetDevice(device);
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};
// device that owns a (device) pointer; the current device if the pointer is unknown to the runtime
static int device_of(const void* ptr) {
    cudaPointerAttributes a;
    if (ptr && cudaPointerGetAttributes(&a, ptr) == cudaSuccess && a.type == cudaMemoryTypeDevice) return a.device;
    cudaGetLastError();
    int d = 0;
    cudaGetDevice(&d);
    return d;
}
struct DeviceState {
    bool ready = false;
    cudaMemPool_t pool = nullptr;
    cudaStream_t upload = nullptr;
};
static DeviceState g_dev[DTC_MAX_DEVICES];
static cudaError_t device_state(int device, DeviceState** out) {       // call with `device` current
    if (device < 0 || device >= DTC_MAX_DEVICES) return cudaErrorInvalidDevice;
    DeviceState& D = g_dev[device];
    if (!D.ready) {
        cudaMemPoolProps props;
        memset(&props, 0, sizeof(props));
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        cudaError_t e = cudaMemPoolCreate(&D.pool, &props);
        if (e != cudaSuccess) return e;
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(D.pool, cudaMemPoolAttrReleaseThreshold, &keep);
        e = cudaStreamCreateWithFlags(&D.upload, cudaStreamNonBlocking);
        if (e != cudaSuccess) return e;
        D.ready = true;
    }
    *out = &D;
    return cudaSuccess;
}
// allocate + fill a table on the device's upload stream (host data is staged before the call returns)
static cudaError_t table_upload(void** ptr, const void* host, size_t bytes, size_t min_bytes, int device) {
    DeviceState* D;
    cudaError_t e = device_state(device, &D);
    if (e != cudaSuccess) return e;
    e = cudaMallocFromPoolAsync(ptr, bytes > min_bytes ? bytes : min_bytes, D->pool, D->upload);
    if (e != cudaSuccess) return e;
    if (bytes) e = cudaMemcpyAsync(*ptr, host, bytes, cudaMemcpyHostToDevice, D->upload);
    return e;
}
static void table_free(void* ptr, int device) {
    if (ptr && device >= 0 && device < DTC_MAX_DEVICES && g_dev[device].ready) cudaFreeAsync(ptr, g_dev[device].upload);
}

struct dtc_program {
    DtcProgramHost h;
    DtcEvent* d_events = nullptr;
    DtcLayer* d_layers = nullptr;
    DtcSmallPlan small;
    long long* d_small_idx = nullptr;
    long long n_small = -1;              // -1: no read-out plan set
    bool fuse_rdm = false;               // dtc_program_set_fused_rdm
    int fused_local_bit = -1;            // tile-local position of the read-out qubit in the last pass (-1: not fusable)
    bool profiling = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaEvent_t> pass_ev;    // profiling: one event after every pass of the last run (per-pass durations)
    int pass_ev_n = 0;
    DtcStreamPass* d_spasses = nullptr;  // pass descriptors on the device (k_tile_resident looks them up per work item)
    bool resident_ok = false;            // every pass runs on the streaming engine and at most two tensor maps are needed
    int tmap_mode[2] = {0, 0}, tmap_g[2] = {0, 0}, n_tmaps = 0;
    bool last_resident = false;
    int last_kernel_launches = 0;
    cudaEvent_t uploaded = nullptr;      // tables are on the device (recorded on the upload stream)
    cudaEvent_t last_use = nullptr;      // end of the last dtc_program_run / dtc_program_readout on the caller's stream
    bool used = false;
    int last_launches = 0;
    bool last_gen_first = false, last_fused = false;     // what the last dtc_program_run did (traffic accounting)
};

// ------------------------------------------------------------------------------------ kernels
// The event list is the same for every trajectory: the block stages it through shared memory in chunks (coalesced loads)
// and each thread walks its own trajectory's frame over the staged chunk.
#define DTC_FRAMES_CHUNK 256
__global__ void k_frames(const DtcEvent* __restrict__ ev, long long n_events, u64* __restrict__ masks,
                         long long n_traj, long long traj_offset, u64 seed, u64* __restrict__ fx,
                         u64* __restrict__ fz, int* __restrict__ ph) {
    __shared__ DtcEvent sev[DTC_FRAMES_CHUNK];
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < n_traj;
    u64 x = 0, z = 0;
    int p = 0;
    for (long long e0 = 0; e0 < n_events; e0 += DTC_FRAMES_CHUNK) {
        const int n = (int)((n_events - e0 < DTC_FRAMES_CHUNK) ? n_events - e0 : DTC_FRAMES_CHUNK);
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(ev + e0);
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(sev);
        for (int i = threadIdx.x; i < n * (int)(sizeof(DtcEvent) / 8); i += blockDim.x) dst[i] = src[i];
        __syncthreads();
        if (live)
            for (int e = 0; e < n; ++e) frame_step(sev[e], (u64)(traj_offset + t), seed, masks + t, n_traj, x, z, p);
        __syncthreads();
    }
    if (live) {
        fx[t] = x;
        fz[t] = z;
        ph[t] = p & 3;
    }
}

__global__ void k_init_basis(double2* __restrict__ state, int n_local, long long n_traj, u64 index) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_traj) state[((u64)t << n_local) + index] = make_double2(1.0, 0.0);
}

template <int S2_LO, bool HAS_X>
__global__ void __launch_bounds__(DTC_THREADS, 3)
k_tile_pass(double2* __restrict__ state, const __grid_constant__ DtcTilePass P,
            const DtcLayer* __restrict__ layers, const u64* __restrict__ masks, long long n_traj,
            u64 rank_bits, int pf_blocks) {
    extern __shared__ __align__(16) unsigned char smraw[];
    TileSmem& sm = *reinterpret_cast<TileSmem*>(smraw);
    const int tid = threadIdx.x;
    const int ntb = P.n_local - DTC_TILE_BITS;
    double2 a[DTC_NREG];
    u64 rmA0;
    {
        // phase 1: coalesced global loads straight into the register file (registers span S1)
        const u64 tile = (u64)blockIdx.x & ((1ull << ntb) - 1);
        const u64 traj = (u64)blockIdx.x >> ntb;
        const u64 base = tile_base_index(tile, P);
        tile_gload(state + (traj << P.n_local), tile_thread_offset(tid, base, P), P, a);
        // warm L2 with the tile a later CTA will load (one lane per 64 B run)
        const u64 b2 = (u64)blockIdx.x + (u64)pf_blocks;
        if (pf_blocks > 0 && b2 < gridDim.x && (tid & 3) == 0) {
            const u64 base2 = tile_base_index(b2 & ((1ull << ntb) - 1), P);
            tile_prefetch(state + ((b2 >> ntb) << P.n_local), tile_thread_offset(tid, base2, P), P);
        }
        const TileMasks M = tile_load_masks(P, masks, n_traj, traj);
        if (tid == 0) {
            sm.base = base;
            sm.rmA = M.rmA;
            sm.rmB = M.rmB;
        }
        rmA0 = M.rmA;
        // diagonal-layer setup overlaps the loads in flight
        if (P.layerD >= 0)
            tile_setup_thread(tid, sm, P, layers[P.layerD], base | (rank_bits << P.n_local), M.m1a, M.m1b, M.m2);
    }
    tile_rot_s1<S2_LO>(a, P.t1, P.tb, rmA0);    // needs only the loads, not the setup
    __syncthreads();
    tile_sm_store13<S2_LO>(tid, sm, a);
    tile_tables_thread<S2_LO>(tid, sm, P);          // after the stores: a[] is dead here
    __syncthreads();

    // phase 2: registers span S2:  R_A|S2, D, R_B|S2
    tile_sm_load2<S2_LO>(tid, sm, a);
    tile_phase2_compute<S2_LO, HAS_X>(tid, a, sm, P, sm.rmA, sm.rmB);
    tile_sm_store2<S2_LO>(tid, sm, a);
    __syncthreads();

    // phase 3: registers span S1 again, coalesced stores
    tile_sm_load13<S2_LO>(tid, sm, a);
    tile_rot_s1<S2_LO>(a, P.t2, P.tb, sm.rmB);
    tile_gstore(state + (((u64)blockIdx.x >> ntb) << P.n_local), tile_thread_offset(tid, sm.base, P), P, a);
}


// ------------------------------------------------------------------------------------ k_tile_stream
// Persistent, warp-specialised version of the fused pass: one CTA per SM, DTC_STREAM_STAGES stage
// buffers of 64 KB.  Warp 8 (one elected lane) drives the TMA engine: bulk / tensor-map loads into a
// free stage (mbarrier full[s]), and -- once a compute warpgroup has signalled done[s] -- the bulk /
// tensor-map store of the finished tile.  Two compute warpgroups (128 threads, tiles k = wg, wg+2, ...)
// run the three in-place phases of dtc_stream.cuh on the stage buffer.  Loads and stores therefore cost
// no registers and no issue slots of the compute warps.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
#ifndef DTC_TRYWAIT_HINT
#define DTC_TRYWAIT_HINT 0      // suspend-time hint (ns) of the try_wait loops; 0: the hardware default
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
#if DTC_TRYWAIT_HINT > 0
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(parity), "r"((uint32_t)DTC_TRYWAIT_HINT) : "memory");
#else
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
#endif
    }
}
__device__ __forceinline__ void wg_barrier(int wg) {
    if (wg == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
    else asm volatile("bar.sync 2, 128;" ::: "memory");
}

template <class PassT>
__device__ __forceinline__ void stream_tma_load(const PassT& P, const CUtensorMap* tmap, const double2* state,
                                                u64 T, uint32_t dst, uint32_t bar) {
    mbar_expect_tx(bar, DTC_TILE * 16);
    if (P.contig) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(state + (T << DTC_TILE_BITS)), "r"(DTC_TILE * 16), "r"(bar) : "memory");
    } else if (P.mode == 3) {
        const int lb = P.g - 7;        // tensor: (256 doubles | bits [7,g) | 32 rows | everything above)
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2,%3,%4,%5}], [%6];"
            ::"r"(dst), "l"(tmap), "r"(0), "r"((int)(T & ((1ull << lb) - 1))), "r"(0), "r"((int)(T >> lb)), "r"(bar)
            : "memory");
    } else {
        const int lb = P.g - 2;
        asm volatile(
            "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2,%3,%4,%5,%6}], [%7];"
            ::"r"(dst), "l"(tmap), "r"(0), "r"((int)(T & ((1ull << lb) - 1))), "r"(0), "r"(0), "r"((int)(T >> lb)), "r"(bar)
            : "memory");
    }
}
template <class PassT>
__device__ __forceinline__ void stream_tma_store(const PassT& P, const CUtensorMap* tmap, double2* state, u64 T,
                                                 uint32_t src) {
    if (P.contig) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     ::"l"(state + (T << DTC_TILE_BITS)), "r"(src), "r"(DTC_TILE * 16) : "memory");
    } else if (P.mode == 3) {
        const int lb = P.g - 7;
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1,%2,%3,%4}], [%5];"
                     ::"l"(tmap), "r"(0), "r"((int)(T & ((1ull << lb) - 1))), "r"(0), "r"((int)(T >> lb)), "r"(src)
                     : "memory");
    } else {
        const int lb = P.g - 2;
        asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%1,%2,%3,%4,%5}], [%6];"
                     ::"l"(tmap), "r"(0), "r"((int)(T & ((1ull << lb) - 1))), "r"(0), "r"(0), "r"((int)(T >> lb)), "r"(src)
                     : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// Optional role timing (build with -DDTC_STREAM_TIMING; tuning only): cycles per role and state, summed over CTAs.
#ifdef DTC_STREAM_TIMING
__device__ unsigned long long g_stream_prof[16];
#define PROF_DECL unsigned long long pf_t = clock64(), pf_acc[4] = {0, 0, 0, 0}
#define PROF_LAP(i) do { const unsigned long long n_ = clock64(); pf_acc[i] += n_ - pf_t; pf_t = n_; } while (0)
#define PROF_FLUSH(base, cond) do { if (cond) for (int i_ = 0; i_ < 4; ++i_) atomicAdd(&g_stream_prof[(base) + i_], pf_acc[i_]); } while (0)
#else
#define PROF_DECL
#define PROF_LAP(i)
#define PROF_FLUSH(base, cond)
#endif

// Phases 1 and 3 are the same code with different tan values; one out-of-line copy keeps the instruction
// footprint of the compute warps small (straight-line code, fetched once per tile per warp).
template <int MODE>
__device__ __noinline__ void stream_phase13_call(int t, double2* tile, double t0, double t1, double t2, double t3, double t4) {
    const double tt[5] = {t0, t1, t2, t3, t4};
    stream_phase13_signed<MODE>(t, tile, tt);
}

template <class PassT>
__device__ __noinline__ void stream_phase2_partial_call(int t, double2* tile, const StreamSlot* tab, const PassT* P,
                                                         u64 rmA, u64 rmB) {
    stream_phase2_partial(t, tile, *tab, *P, rmA, rmB);
}

struct StreamSmem {
    double2 stage[DTC_STREAM_STAGES][DTC_TILE];
    StreamSlot slot[DTC_STREAM_STAGES];
    StreamBuild build[DTC_STREAM_STAGES];
    DtcLayer layer;                                   // D layer of this pass (fixed for the whole launch)
    unsigned long long full[DTC_STREAM_STAGES], done[DTC_STREAM_STAGES];
};
static_assert(sizeof(StreamSmem) + 128 <= 227 * 1024, "stage buffers + tables must fit one CTA's shared memory");

// Warp roles: warps 0..7 two compute warpgroups; warp 8 TMA driver (one lane); warps 9.. one table builder per stage.
// full[s] completes when the tile's bytes have landed AND its tables are written (2 arrivals + tx bytes);
// done[s] completes when all 128 threads of the warpgroup have finished the tile in stage s.
template <int MODE>
__global__ void __launch_bounds__(DTC_STREAM_THREADS, 1)
k_tile_stream(double2* __restrict__ state, const __grid_constant__ CUtensorMap tmap, const __grid_constant__ DtcStreamPass P,
              const DtcLayer* __restrict__ layers, const u64* __restrict__ masks, long long n_traj, u64 rank_bits,
              long long n_tiles, u64 init_index, double2* __restrict__ rdm_out, int rdm_local_bit, double2* store_base,
              int store_lag) {
    // store_base != state (contiguous tiles only): the finished tiles are stored THERE instead of in place -- e.g. straight
    // into a peer GPU's receive buffer over NVLink, fusing the last sweep before an exchange with the exchange itself.
    // store_lag = 1 (for such slow stores): a stage is reloaded only after the NEXT tile's store has been issued, so two
    // stores per SM are in flight instead of one (the loads run one tile later; a store-bound sweep does not care).
    // rdm_out != nullptr (last pass of a factorised circuit): the tile is not stored; the reduced density matrix of the
    // tile-local bit rdm_local_bit of psi' is accumulated into rdm_out[trajectory][2][2] instead.
    // init_index != DTC_INIT_KEEP: the input is not read -- every trajectory starts in the basis state init_index
    // (DTC_INIT_ZERO: the zero vector).  Tiles without the basis amplitude are written as zeros without compute.
    const bool gen = init_index != DTC_INIT_KEEP;
    extern __shared__ unsigned char smraw[];
    StreamSmem& sm = *reinterpret_cast<StreamSmem*>(smraw + ((128u - (smem_u32(smraw) & 127u)) & 127u));
    const int tid = threadIdx.x, warp = tid >> 5;
    const long long K = (n_tiles - (long long)blockIdx.x + (long long)gridDim.x - 1) / (long long)gridDim.x;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < DTC_STREAM_STAGES; ++s) {
            mbar_init(smem_u32(&sm.full[s]), 2);
            mbar_init(smem_u32(&sm.done[s]), 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (P.layerD >= 0) {
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(layers + P.layerD);
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(&sm.layer);
        for (int i = tid; i < (int)(sizeof(DtcLayer) / 8); i += DTC_STREAM_THREADS) dst[i] = src[i];
    }
    __syncthreads();
    if (warp == 4 * DTC_STREAM_WG) {
        // ---- TMA driver
        if ((tid & 31) != 0) return;
        for (long long k = 0; k < K && k < DTC_STREAM_STAGES; ++k) {
            if (gen) mbar_arrive(smem_u32(&sm.full[k]));
            else stream_tma_load(P, &tmap, state, (u64)blockIdx.x + (u64)k * gridDim.x, smem_u32(sm.stage[k]), smem_u32(&sm.full[k]));
        }
        int s = 0;
        uint32_t par = 0;
        PROF_DECL;
        for (long long k = 0; k < K; ++k) {
            mbar_wait(smem_u32(&sm.done[s]), par);
            PROF_LAP(0);
            if (!rdm_out) stream_tma_store(P, &tmap, store_base, (u64)blockIdx.x + (u64)k * gridDim.x, smem_u32(sm.stage[s]));
            if (store_lag) {
                // reload the stage of the PREVIOUS tile: all but the store just issued have read their stage
                if (k >= 1 && k - 1 + DTC_STREAM_STAGES < K) {
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    const int sp = s == 0 ? DTC_STREAM_STAGES - 1 : s - 1;
                    stream_tma_load(P, &tmap, state, (u64)blockIdx.x + (u64)(k - 1 + DTC_STREAM_STAGES) * gridDim.x,
                                    smem_u32(sm.stage[sp]), smem_u32(&sm.full[sp]));
                }
            } else if (k + DTC_STREAM_STAGES < K) {
                if (!rdm_out) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the store has read the stage
                PROF_LAP(1);
                if (gen) mbar_arrive(smem_u32(&sm.full[s]));
                else stream_tma_load(P, &tmap, state, (u64)blockIdx.x + (u64)(k + DTC_STREAM_STAGES) * gridDim.x,
                                     smem_u32(sm.stage[s]), smem_u32(&sm.full[s]));
            }
            if (++s == DTC_STREAM_STAGES) { s = 0; par ^= 1u; }
            PROF_LAP(2);
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        PROF_LAP(3);
        PROF_FLUSH(0, true);
        return;
    }
    const int ntb = P.n_local - DTC_TILE_BITS;
    if (warp > 4 * DTC_STREAM_WG) {
        // ---- table builder of stage s: phase tables + rotation sign masks of tiles k = s, s + stages, ...
        const int lane = tid & 31, s = warp - (4 * DTC_STREAM_WG + 1);
        StreamBuild& bl = sm.build[s];
        StreamSlot& slot = sm.slot[s];
        if (s >= K) return;
        StreamMasks M = stream_load_masks(P, masks, n_traj, ((u64)blockIdx.x + (u64)s * gridDim.x) >> ntb);
        uint32_t par = 1;                                  // parity of done[s] for the PREVIOUS use of the slot
        PROF_DECL;
        for (long long k = s; k < K; k += DTC_STREAM_STAGES) {
            const u64 T = (u64)blockIdx.x + (u64)k * gridDim.x;
            // next tile's masks: issued now, consumed in the next iteration
            const u64 Tn = (k + DTC_STREAM_STAGES < K) ? T + (u64)DTC_STREAM_STAGES * gridDim.x : T;
            const StreamMasks Mn = stream_load_masks(P, masks, n_traj, Tn >> ntb);
            const u64 base = stream_tile_base(T & ((1ull << ntb) - 1), P);
            if (P.layerD >= 0) {
                stream_build1(lane, bl, P, sm.layer, base | (rank_bits << P.n_local), M.m1a, M.m1b, M.m2);
                __syncwarp();
                stream_build2(lane, bl, P, sm.layer);
                __syncwarp();
            }
            PROF_LAP(0);
            if (k >= DTC_STREAM_STAGES) mbar_wait(smem_u32(&sm.done[s]), par);      // slot s is free again
            PROF_LAP(1);
            stream_build3(lane, bl, slot, P);
            if (lane == 0) { slot.rmA = M.rmA; slot.rmB = M.rmB; }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&sm.full[s]));
            M = Mn;
            par ^= 1u;
            PROF_LAP(2);
        }
        PROF_FLUSH(4, lane == 0 && s == 0);
        return;
    }
    // ---- compute warpgroups
    const int wg = warp >> 2, t = tid & 127;

    PROF_DECL;
    for (long long k = wg; k < K; k += DTC_STREAM_WG) {
        const int s = (int)(k % DTC_STREAM_STAGES);
        const uint32_t u = (uint32_t)(k / DTC_STREAM_STAGES);
        if (u > 0) mbar_wait(smem_u32(&sm.done[s]), (u - 1) & 1u);     // never run a full phase ahead of the stage
        mbar_wait(smem_u32(&sm.full[s]), u & 1u);
        // mode A: a warp owns local bits 10,11 in all phases, so its quarter of the tile is private and __syncwarp()
        // is all the phases need (measured 1-2 % faster than warpgroup barriers once the code footprint was small);
        // mode B exchanges data between the warps of the warpgroup
        if (MODE != 1) wg_barrier(wg);
        PROF_LAP(0);
        double2* tile = sm.stage[s];
        const StreamSlot& slot = sm.slot[s];
        const u64 rmA = slot.rmA, rmB = slot.rmB;
        if (gen) {
            const u64 T = (u64)blockIdx.x + (u64)k * gridDim.x;
            const u64 base = stream_tile_base(T & ((1ull << ntb) - 1), P);
#pragma unroll
            for (int r = 0; r < DTC_NREG; ++r) tile[t + 128 * r] = make_double2(0.0, 0.0);
            const bool has = init_index != DTC_INIT_ZERO && ((init_index ^ base) & ~P.tile_mask) == 0;
            if (!has) {                                    // all-zero tile: nothing to compute
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(smem_u32(&sm.done[s]));
                continue;
            }
            wg_barrier(wg);
            if (t == 0) {
                int l = 0;
                for (int b = 0; b < DTC_TILE_BITS; ++b) l |= (int)((init_index >> P.tb[b]) & 1ull) << b;
                tile[l] = make_double2(1.0, 0.0);
            }
            wg_barrier(wg);
        }
        if (MODE == 3) {
            // one register set, one phase: thread <-> the seven passive low bits, registers <-> the five active bits
            stream_phaseC(t, tile, slot, P, rmA, rmB);
            PROF_LAP(2);
        } else {
            constexpr int M13 = MODE == 3 ? 1 : MODE;
            double tt[5];
            if (P.layerA >= 0) {                   // no rotations of layer j in this pass: nothing to do on S1 before the diagonal
                stream_signed_s1<M13>(P.t1, P.tb, rmA, tt);
                stream_phase13_call<M13>(t, tile, tt[0], tt[1], tt[2], tt[3], tt[4]);
            }
            if (MODE == 1) __syncwarp(); else wg_barrier(wg);
            PROF_LAP(1);
            if (P.layerA >= 0 && P.layerD >= 0 && P.layerB >= 0) stream_phase2(t, tile, slot, P, rmA, rmB);
            else stream_phase2_partial_call(t, tile, &slot, &P, rmA, rmB);
            if (MODE == 1) __syncwarp(); else wg_barrier(wg);
            PROF_LAP(2);
            if (P.layerB >= 0) {
                stream_signed_s1<M13>(P.t2, P.tb, rmB, tt);
                stream_phase13_call<M13>(t, tile, tt[0], tt[1], tt[2], tt[3], tt[4]);
            }
        }
        if (rdm_out) {
            wg_barrier(wg);
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            stream_rdm_pairs(t, tile, rdm_local_bit, acc);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
            if ((t & 31) == 0) {
                const u64 T = (u64)blockIdx.x + (u64)k * gridDim.x;
                double* dst = reinterpret_cast<double*>(rdm_out + ((T >> ntb) << 2));
                atomicAdd(dst + 0, acc[0]);                 // rho[0][0]
                atomicAdd(dst + 2, acc[2]);                 // rho[0][1] = sum a conj(b)
                atomicAdd(dst + 3, acc[3]);
                atomicAdd(dst + 4, acc[2]);                 // rho[1][0] = conj
                atomicAdd(dst + 5, -acc[3]);
                atomicAdd(dst + 6, acc[1]);                 // rho[1][1]
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the TMA store
        mbar_arrive(smem_u32(&sm.done[s]));
        PROF_LAP(3);
    }
    PROF_FLUSH(8, t == 0 && wg == 0);
    PROF_FLUSH(12, t == 0 && wg == 1);
}


// ------------------------------------------------------------------------------------ k_tile_resident
// The whole pass schedule of a circuit in ONE persistent launch, ordered so that the states stay in L2 (126 MB): the
// trajectories are taken in groups of G (G x 2^n_local x 16 B ~ 64 MiB), a group runs through ALL passes before the next
// group starts, and every group reuses the same G state slots -- the sweeps read and write L2, not HBM.  Work items
// (group, pass, slot, tile) are numbered in that order and dealt to the CTAs round robin; an item may be loaded once every
// tile of the same slot in the previous pass (for pass 0: in the last pass of the previous group) has been stored, which
// per-(pass, slot) completion counters in global memory track.  The dependency of an item lies G x tiles-per-state items
// earlier in the order, further than the CTAs' pipelines reach, so the counters are polled but rarely waited on.
// Roles, stage ring, tables and phases are those of k_tile_stream; the pass descriptor is looked up per item.
struct ResidentPlan {
    int n_local, n_passes, G, last_G;      // G: state slots = trajectories per group (last_G: of the last group)
    int nt_bits;                            // tiles per state = 1 << nt_bits
    int gen_first, fused_last, rdm_local_bit;
    long long n_traj, n_groups, items_per_group, total_items;
    u64 init_index, rank_bits;
};
struct ResidentItem {
    int p, j;                               // pass, state slot
    u64 T, traj;                            // tile within the state, trajectory within the batch
    long long seq;                          // group * n_passes + pass
};
// Iterator over the work items w0, w0 + step, w0 + 2 step, ... of one role: (group, offset in group) advanced without divisions
struct ResidentIter {
    long long gi;                           // group
    unsigned r;                             // offset inside the group
    unsigned step;
    __device__ __forceinline__ void init(long long w0, unsigned step_, const ResidentPlan& R) {
        gi = w0 / R.items_per_group;
        r = (unsigned)(w0 - gi * R.items_per_group);
        step = step_;
    }
    __device__ __forceinline__ void next(const ResidentPlan& R) {
        r += step;
        const unsigned ipg = (unsigned)R.items_per_group;
        while (r >= ipg && gi < R.n_groups - 1) { r -= ipg; ++gi; }
    }
    __device__ __forceinline__ ResidentItem item(const ResidentPlan& R) const {
        ResidentItem it;
        const int Gi = (gi == R.n_groups - 1) ? R.last_G : R.G;
        const unsigned per_pass = (unsigned)Gi << R.nt_bits;
        it.p = (int)(r / per_pass);
        const unsigned rr = r - (unsigned)it.p * per_pass;
        it.j = (int)(rr >> R.nt_bits);
        it.T = (u64)(rr & ((1u << R.nt_bits) - 1));
        it.traj = (u64)(gi * R.G + it.j);
        it.seq = gi * R.n_passes + it.p;
        return it;
    }
};
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
// Completion counters live in L2 (the point of coherence of the TMA loads and stores they order): a relaxed gpu-scope load
// observes them there, and the writer counts an item with a RED only after its bulk stores have completed.
__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Issued only after cp.async.bulk.wait_group has reported the item's bulk stores complete, i.e. landed in L2.
__device__ __forceinline__ void red_relaxed_gpu(int* p) {
    asm volatile("red.relaxed.gpu.global.add.s32 [%0], 1;" ::"l"(p) : "memory");
}

struct ResidentPasses {
    DtcResidentPass p[DTC_RESIDENT_MAX_PASSES];
};
static_assert(sizeof(ResidentPasses) + sizeof(ResidentPlan) + 2 * sizeof(CUtensorMap) + 64 <= 32764, "kernel parameter space");
static_assert(offsetof(DtcLayer, rot_any) == DTC_LAYER_PREFIX, "DTC_LAYER_PREFIX must cover what the table builders read");

// what the TMA driver needs of a pass
struct ResidentTma {
    int contig, mode, g;
};
struct ResidentPost {
    u64 tile;                               // tile coordinate in the slot buffer: (slot << nt_bits) | tile in state
    int counter;                            // index of the item's completion counter
    unsigned char contig, mode, g, tmap;    // of its pass
    unsigned char gen, readonly, pad0, pad1;   // first pass generates the state (no load); fused last pass (no store)
};

struct ResidentSmem {
    double2 stage[DTC_STREAM_STAGES][DTC_TILE];
    StreamSlot slot[DTC_STREAM_STAGES];
    StreamBuild build[DTC_STREAM_STAGES];
    // private to the stage's table-builder warp: the full descriptor and the layer tables of the pass it is building for,
    // refreshed when the pass changes (global reads would miss the few KB of L1 left beside 217 KB of shared memory)
    DtcStreamPass bpass[DTC_STREAM_STAGES];
    unsigned long long blayer[DTC_STREAM_STAGES][DTC_LAYER_PREFIX / 8];
    unsigned long long full[DTC_STREAM_STAGES], done[DTC_STREAM_STAGES];
    long long dep_ok[DTC_STREAM_STAGES];
    ResidentPlan plan;                                 // copy of the launch plan for the out-of-line roles
    unsigned char pinfo[DTC_RESIDENT_MAX_PASSES][4];   // per pass: contig, mode, g, tensor-map slot (TMA driver)
    // what the TMA driver needs of an item, posted by the stage's builder (the driver lane is a single thread on the critical
    // path of every tile: it does no index arithmetic of its own).  Three deep per stage: an entry is overwritten only
    // after two later items of the stage have been loaded, i.e. long after its own store was issued.
    ResidentPost post[DTC_STREAM_STAGES][3];
    int pend[4];
};
static_assert(sizeof(ResidentSmem) + 128 <= 227 * 1024, "stage buffers + tables must fit one CTA's shared memory");

// ---- TMA driver (one lane).  Out of line: the role gets a register allocation of its own -- inlined into the kernel its loop
// state is spilled to local memory (the compute warps need every register), and with 217 KB of shared memory per CTA there is
// next to no L1 to catch those spills.
// An event loop that never blocks on another CTA while it has a store to issue.  Per item it issues the store, waits until the
// TMA engine has read the stage (the one long wait, as in k_tile_stream) and reloads the stage.  Whether the reload's
// dependency is met it reads from shared memory: the stage's table-builder warp, which runs ahead and has slack, polls the
// completion counter in global memory and posts the result.  Finished stores are counted when the loop is idle.
__device__ __noinline__ void resident_driver(ResidentSmem* smp, double2* state, const CUtensorMap* tmap0, const CUtensorMap* tmap1,
                                             int* cnt, long long K) {
    ResidentSmem& sm = *smp;
    long long load_k = 0, store_k = 0;
    int ls = 0, lu = 0, ss = 0, su = 0;               // stage and use count (mod 3) of the load / store candidate
    uint32_t spar = 0;                                // parity of done[ss] for the store candidate
    int* pend = sm.pend;                              // counters of items stored but not yet counted
    int n_pend = 0;
    PROF_DECL;
    while (store_k < K) {
        bool progressed = false;
        PROF_LAP(0);
        if (store_k < load_k && mbar_test(smem_u32(&sm.done[ss]), spar)) {
            const ResidentPost it = sm.post[ss][su];
            if (it.readonly) {
                red_relaxed_gpu(cnt + it.counter);    // read-only item: complete as soon as the stage has been consumed
            } else {
                if (n_pend == 4) {
                    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                    for (int i = 0; i < n_pend; ++i) red_relaxed_gpu(cnt + pend[i]);
                    n_pend = 0;
                }
                const ResidentTma P = {it.contig, it.mode, it.g};
                stream_tma_store(P, it.tmap ? tmap1 : tmap0, state, it.tile, smem_u32(sm.stage[ss]));
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the stage may be reloaded
                pend[n_pend++] = it.counter;
            }
            ++store_k;
            if (++ss == DTC_STREAM_STAGES) { ss = 0; spar ^= 1u; if (++su == 3) su = 0; }
            progressed = true;
            PROF_LAP(1);
        }
        if (load_k < K && load_k < store_k + DTC_STREAM_STAGES) {
            // dep_ok[s] = 1 + (index of the newest item of stage s whose dependency the builder has seen met and whose
            // description it has posted)
            if (*(volatile long long*)&sm.dep_ok[ls] > load_k) {
                const ResidentPost it = sm.post[ls][lu];
                if (it.gen) {
                    mbar_arrive(smem_u32(&sm.full[ls]));
                } else {
                    // (no proxy fence: the load is issued only after the counter value has been observed -- a control
                    //  dependency -- and it reads L2, where the producers' bulk stores had landed before they counted)
                    const ResidentTma P = {it.contig, it.mode, it.g};
                    stream_tma_load(P, it.tmap ? tmap1 : tmap0, state, it.tile, smem_u32(sm.stage[ls]), smem_u32(&sm.full[ls]));
                }
                ++load_k;
                if (++ls == DTC_STREAM_STAGES) { ls = 0; if (++lu == 3) lu = 0; }
                progressed = true;
                PROF_LAP(2);
            }
        }
        if (!progressed) {
            if (n_pend) {                              // idle: count what has been stored (bounded wait on the TMA engine only)
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                for (int i = 0; i < n_pend; ++i) red_relaxed_gpu(cnt + pend[i]);
                n_pend = 0;
            } else if (load_k == K || load_k == store_k + DTC_STREAM_STAGES) {
                // every stage is occupied: the only thing that can happen next is the oldest tile being finished --
                // a hardware-assisted wait on its barrier
                mbar_wait(smem_u32(&sm.done[ss]), spar);
            }
            // else: a free stage waits for its dependency (posted by the builder): poll again at once
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    for (int i = 0; i < n_pend; ++i) red_relaxed_gpu(cnt + pend[i]);
    PROF_LAP(3);
    PROF_FLUSH(0, true);
}

// ---- table builder of stage s (one warp), out of line for the same reason
__device__ __noinline__ void resident_builder(ResidentSmem* smp, int s, int lane, const DtcStreamPass* __restrict__ passes,
                                              const DtcLayer* __restrict__ layers, const u64* __restrict__ masks,
                                              const int* cnt, long long K) {
    ResidentSmem& sm = *smp;
    const ResidentPlan& R = sm.plan;
    StreamBuild& bl = sm.build[s];
    StreamSlot& slot = sm.slot[s];
    const int nt = 1 << R.nt_bits;
    uint32_t par = 1;
    ResidentIter bi;
    bi.init((long long)blockIdx.x + (long long)s * gridDim.x, DTC_STREAM_STAGES * gridDim.x, R);
    PROF_DECL;
    DtcStreamPass& P = sm.bpass[s];
    const DtcLayer& L = *reinterpret_cast<const DtcLayer*>(sm.blayer[s]);
    int cached_p = -1;
    for (long long k = s; k < K; k += DTC_STREAM_STAGES, bi.next(R)) {
        const ResidentItem it = bi.item(R);
        if (it.p != cached_p) {                       // new pass: stage its descriptor and layer tables (warp-private)
            const uint4* src = reinterpret_cast<const uint4*>(passes + it.p);
            uint4* dst = reinterpret_cast<uint4*>(&sm.bpass[s]);
            for (int i = lane; i < (int)(sizeof(DtcStreamPass) / 16); i += 32) dst[i] = src[i];
            __syncwarp();
            if (P.layerD >= 0) {
                const uint2* ls = reinterpret_cast<const uint2*>(layers + P.layerD);
                uint2* ld = reinterpret_cast<uint2*>(sm.blayer[s]);
                for (int i = lane; i < DTC_LAYER_PREFIX / 8; i += 32) ld[i] = ls[i];
            }
            __syncwarp();
            cached_p = it.p;
        }
        const StreamMasks M = stream_load_masks(P, masks, R.n_traj, it.traj);
        const u64 base = stream_tile_base(it.T, P);
        if (P.layerD >= 0) {
            stream_build1(lane, bl, P, L, base | (R.rank_bits << P.n_local), M.m1a, M.m1b, M.m2);
            __syncwarp();
            stream_build2(lane, bl, P, L);
            __syncwarp();
        }
        PROF_LAP(0);
        if (lane == 0) {
            ResidentPost po;
            po.tile = ((u64)it.j << R.nt_bits) | it.T;
            po.counter = (int)(it.seq * R.G + it.j);
            po.contig = (unsigned char)P.contig; po.mode = (unsigned char)P.mode; po.g = (unsigned char)P.g;
            po.tmap = (unsigned char)P.tmap_slot;
            po.gen = (it.p == 0 && R.gen_first) ? 1 : 0;
            po.readonly = (R.fused_last && it.p == R.n_passes - 1) ? 1 : 0;
            po.pad0 = po.pad1 = 0;
            sm.post[s][(k / DTC_STREAM_STAGES) % 3] = po;
            // the item's dependency: every tile of its state slot in the previous pass (previous group's last pass for
            // pass 0) has been stored.  Polled here, ahead of time and off the TMA driver's critical path; the driver
            // reads the outcome from shared memory.
            if (it.seq > 0) {
                const int* c = cnt + (it.seq - 1) * R.G + it.j;
                while (ld_relaxed_gpu(c) < nt) __nanosleep(64);
            }
            __threadfence_block();
            *(volatile long long*)&sm.dep_ok[s] = k + 1;
        }
        PROF_LAP(3);
        if (k >= DTC_STREAM_STAGES) mbar_wait(smem_u32(&sm.done[s]), par);      // slot s is free again
        PROF_LAP(1);
        stream_build3(lane, bl, slot, P);
        if (lane == 0) { slot.rmA = M.rmA; slot.rmB = M.rmB; }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&sm.full[s]));
        par ^= 1u;
        PROF_LAP(2);
    }
    PROF_FLUSH(4, lane == 0 && s == 0);
}

__global__ void __launch_bounds__(DTC_STREAM_THREADS, 1)
k_tile_resident(double2* __restrict__ state, const __grid_constant__ CUtensorMap tmap0, const __grid_constant__ CUtensorMap tmap1,
                const __grid_constant__ ResidentPlan R, const __grid_constant__ ResidentPasses RP,
                const DtcStreamPass* __restrict__ passes, const DtcLayer* __restrict__ layers,
                const u64* __restrict__ masks, int* __restrict__ cnt, double2* __restrict__ rdm_out) {
    extern __shared__ unsigned char smraw[];
    ResidentSmem& sm = *reinterpret_cast<ResidentSmem*>(smraw + ((128u - (smem_u32(smraw) & 127u)) & 127u));
    const int tid = threadIdx.x, warp = tid >> 5;
    const long long K = (R.total_items - (long long)blockIdx.x + (long long)gridDim.x - 1) / (long long)gridDim.x;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < DTC_STREAM_STAGES; ++s) {
            mbar_init(smem_u32(&sm.full[s]), 2);
            mbar_init(smem_u32(&sm.done[s]), 128);
            sm.dep_ok[s] = 0;
        }
        sm.plan = R;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < R.n_passes) {
        sm.pinfo[tid][0] = RP.p[tid].contig;
        sm.pinfo[tid][1] = RP.p[tid].mode;
        sm.pinfo[tid][2] = RP.p[tid].g;
        sm.pinfo[tid][3] = RP.p[tid].tmap_slot;
    }
    __syncthreads();
    if (warp == 4 * DTC_STREAM_WG) {
        if ((tid & 31) == 0) resident_driver(&sm, state, &tmap0, &tmap1, cnt, K);
        return;
    }
    if (warp > 4 * DTC_STREAM_WG) {
        const int s = warp - (4 * DTC_STREAM_WG + 1);
        if (s < K) resident_builder(&sm, s, tid & 31, passes, layers, masks, cnt, K);
        return;
    }
    // ---- compute warpgroups
    const int wg = warp >> 2, t = tid & 127;
    ResidentIter ci;
    ci.init((long long)blockIdx.x + (long long)wg * gridDim.x, DTC_STREAM_WG * gridDim.x, R);
    PROF_DECL;
    for (long long k = wg; k < K; k += DTC_STREAM_WG, ci.next(R)) {
        const int s = (int)(k % DTC_STREAM_STAGES);
        const uint32_t u = (uint32_t)(k / DTC_STREAM_STAGES);
        const ResidentItem it = ci.item(R);
        const DtcResidentPass& P = RP.p[it.p];
        const int mode = P.mode;
        if (u > 0) mbar_wait(smem_u32(&sm.done[s]), (u - 1) & 1u);
        mbar_wait(smem_u32(&sm.full[s]), u & 1u);
        if (mode != 1) wg_barrier(wg);
        PROF_LAP(0);
        double2* tile = sm.stage[s];
        const StreamSlot& slot = sm.slot[s];
        const u64 rmA = slot.rmA, rmB = slot.rmB;
        if (it.p == 0 && R.gen_first) {
            const u64 base = stream_tile_base(it.T, P);
#pragma unroll
            for (int r = 0; r < DTC_NREG; ++r) tile[t + 128 * r] = make_double2(0.0, 0.0);
            const bool has = R.init_index != DTC_INIT_ZERO && ((R.init_index ^ base) & ~P.tile_mask) == 0;
            if (!has) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(smem_u32(&sm.done[s]));
                continue;
            }
            wg_barrier(wg);
            if (t == 0) {
                int l = 0;
                for (int b = 0; b < DTC_TILE_BITS; ++b) l |= (int)((R.init_index >> P.tb[b]) & 1ull) << b;
                tile[l] = make_double2(1.0, 0.0);
            }
            wg_barrier(wg);
        }
        if (mode == 3) {
            stream_phaseC(t, tile, slot, P, rmA, rmB);
        } else {
            double tt[5];
            if (P.layerA >= 0) {
                if (mode == 1) {
                    stream_signed_s1<1>(P.t1, P.tb, rmA, tt);
                    stream_phase13_call<1>(t, tile, tt[0], tt[1], tt[2], tt[3], tt[4]);
                } else {
                    stream_signed_s1<2>(P.t1, P.tb, rmA, tt);
                    stream_phase13_call<2>(t, tile, tt[0], tt[1], tt[2], tt[3], tt[4]);
                }
            }
            if (mode == 1) __syncwarp(); else wg_barrier(wg);
            PROF_LAP(1);
            if (P.layerA >= 0 && P.layerD >= 0 && P.layerB >= 0) stream_phase2(t, tile, slot, P, rmA, rmB);
            else stream_phase2_partial_call(t, tile, &slot, &P, rmA, rmB);
            if (mode == 1) __syncwarp(); else wg_barrier(wg);
            PROF_LAP(2);
            if (P.layerB >= 0) {
                if (mode == 1) {
                    stream_signed_s1<1>(P.t2, P.tb, rmB, tt);
                    stream_phase13_call<1>(t, tile, tt[0], tt[1], tt[2], tt[3], tt[4]);
                } else {
                    stream_signed_s1<2>(P.t2, P.tb, rmB, tt);
                    stream_phase13_call<2>(t, tile, tt[0], tt[1], tt[2], tt[3], tt[4]);
                }
            }
        }
        if (R.fused_last && it.p == R.n_passes - 1) {
            wg_barrier(wg);
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            stream_rdm_pairs(t, tile, R.rdm_local_bit, acc);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
            if ((t & 31) == 0) {
                double* dst = reinterpret_cast<double*>(rdm_out + (it.traj << 2));
                atomicAdd(dst + 0, acc[0]);
                atomicAdd(dst + 2, acc[2]);
                atomicAdd(dst + 3, acc[3]);
                atomicAdd(dst + 4, acc[2]);
                atomicAdd(dst + 5, -acc[3]);
                atomicAdd(dst + 6, acc[1]);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(smem_u32(&sm.done[s]));
        PROF_LAP(3);
    }
    PROF_FLUSH(8, t == 0 && wg == 0);
    PROF_FLUSH(12, t == 0 && wg == 1);
}

// ---- generic engine
__global__ void k_generic_rot(double2* __restrict__ state, int n_local, int q, double t,
                              const u64* __restrict__ rmask, long long n_traj) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long npairs = n_traj << (n_local - 1);
    if (i >= npairs) return;
    const u64 traj = (u64)i >> (n_local - 1);
    const u64 p = (u64)i & ((1ull << (n_local - 1)) - 1);
    const u64 lowm = (1ull << q) - 1;
    const u64 i0 = ((p & ~lowm) << 1) | (p & lowm);
    double2* st = state + (traj << n_local);
    const double ts = ((rmask[traj] >> q) & 1ull) ? -t : t;
    double2 x0 = st[i0], x1 = st[i0 | (1ull << q)];
    rot_pair(x0, x1, ts);
    st[i0] = x0;
    st[i0 | (1ull << q)] = x1;
}

__global__ void k_generic_diag(double2* __restrict__ state, int n_local, const DtcLayer* __restrict__ L,
                               const u64* __restrict__ masks, long long n_traj, u64 rank_bits) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (n_traj << n_local)) return;
    const u64 traj = (u64)i >> n_local;
    const u64 x = (u64)i & ((1ull << n_local) - 1);
    const u64 m1a = masks[1 * n_traj + traj], m1b = masks[2 * n_traj + traj], m2 = masks[3 * n_traj + traj];
    const double2 ph = diag_phase(*L, x | (rank_bits << n_local), m1a, m1b, m2);
    state[i] = cmul(state[i], ph);
}

// ---- read-out
__global__ void k_materialize(double2* __restrict__ state, int n_local, long long n_traj,
                              const u64* __restrict__ fx, const u64* __restrict__ fz, const int* __restrict__ ph) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (n_traj << n_local)) return;
    const u64 traj = (u64)i >> n_local;
    const u64 y = (u64)i & ((1ull << n_local) - 1);
    const u64 X = fx[traj], Z = fz[traj];
    const u64 y2 = y ^ X;
    if (X != 0 && y > y2) return;                 // the pair is handled by its smaller member
    double2* st = state + (traj << n_local);
    const int p = ph[traj] & 3;
    auto phase = [p](double2 v, int neg) {
        double2 r;
        switch (p) {
            case 0: r = v; break;
            case 1: r = make_double2(-v.y, v.x); break;
            case 2: r = make_double2(-v.x, -v.y); break;
            default: r = make_double2(v.y, -v.x); break;
        }
        if (neg) { r.x = -r.x; r.y = -r.y; }
        return r;
    };
    // psi_true(y) = i^ph (-1)^{popc((y^X)&Z)} psi'(y^X)
    const double2 vy = st[y], vy2 = st[y2];
    const double2 ny = phase(vy2, __popcll(y2 & Z) & 1);
    st[y] = ny;
    if (X != 0) st[y2] = phase(vy, __popcll(y & Z) & 1);
}

__global__ void k_probs(const double2* __restrict__ state, int n_local, long long n_traj, int k,
                        const int* __restrict__ qubits, const u64* __restrict__ fx, double* __restrict__ out,
                        int chunks_per_traj) {
    extern __shared__ double sbins[];
    const int nb = 1 << k;
    for (int b = threadIdx.x; b < nb; b += blockDim.x) sbins[b] = 0.0;
    __syncthreads();
    const u64 traj = blockIdx.x / chunks_per_traj;
    const u64 chunk = blockIdx.x % chunks_per_traj;
    const u64 per = (1ull << n_local) / chunks_per_traj;
    const double2* st = state + (traj << n_local);
    int flip = 0;
    if (fx) {
        const u64 X = fx[traj];
        for (int b = 0; b < k; ++b) flip |= (int)((X >> qubits[b]) & 1ull) << b;
    }
    if (k <= 2) {
        double acc[4] = {0, 0, 0, 0};
        for (u64 x = chunk * per + threadIdx.x; x < (chunk + 1) * per; x += blockDim.x) {
            const double2 v = st[x];
            const double p = v.x * v.x + v.y * v.y;
            int bin = 0;
            for (int b = 0; b < k; ++b) bin |= (int)((x >> qubits[b]) & 1ull) << b;
            bin ^= flip;
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[c] += (bin == c) ? p : 0.0;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            double v = acc[c];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0 && c < nb) atomicAdd(&sbins[c], v);
        }
    } else {
        for (u64 x = chunk * per + threadIdx.x; x < (chunk + 1) * per; x += blockDim.x) {
            const double2 v = st[x];
            int bin = 0;
            for (int b = 0; b < k; ++b) bin |= (int)((x >> qubits[b]) & 1ull) << b;
            atomicAdd(&sbins[bin ^ flip], v.x * v.x + v.y * v.y);
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nb; b += blockDim.x) atomicAdd(&out[traj * nb + b], sbins[b]);
}

__global__ void k_expect_z(const double2* __restrict__ state, int n_local, long long n_traj,
                           const u64* __restrict__ fx, double* __restrict__ out, int chunks_per_traj) {
    __shared__ double sacc[DTC_MAXQ];
    if (threadIdx.x < DTC_MAXQ) sacc[threadIdx.x] = 0.0;
    __syncthreads();
    const u64 traj = blockIdx.x / chunks_per_traj;
    const u64 chunk = blockIdx.x % chunks_per_traj;
    const u64 per = (1ull << n_local) / chunks_per_traj;
    const double2* st = state + (traj << n_local);
    double acc[DTC_MAXQ];
#pragma unroll
    for (int q = 0; q < DTC_MAXQ; ++q) acc[q] = 0.0;
    for (u64 x = chunk * per + threadIdx.x; x < (chunk + 1) * per; x += blockDim.x) {
        const double2 v = st[x];
        const double p = v.x * v.x + v.y * v.y;
#pragma unroll
        for (int q = 0; q < DTC_MAXQ; ++q)
            if (q < n_local) acc[q] += ((x >> q) & 1ull) ? -p : p;
    }
    const u64 X = fx ? fx[traj] : 0ull;
#pragma unroll
    for (int q = 0; q < DTC_MAXQ; ++q) {
        if (q < n_local) {
            double v = acc[q];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0) atomicAdd(&sacc[q], ((X >> q) & 1ull) ? -v : v);
        }
    }
    __syncthreads();
    if (threadIdx.x < n_local) atomicAdd(&out[traj * n_local + threadIdx.x], sacc[threadIdx.x]);
}

template <int K>
__global__ void k_rdm(const double2* __restrict__ state, int n_local, long long n_traj, const int* __restrict__ qubits,
                      double2* __restrict__ out, int chunks_per_traj) {
    constexpr int D = 1 << K;
    const u64 traj = blockIdx.x / chunks_per_traj;
    const u64 chunk = blockIdx.x % chunks_per_traj;
    const u64 nrest = 1ull << (n_local - K);
    const u64 per = nrest / chunks_per_traj;
    const double2* st = state + (traj << n_local);
    int qs[K > 0 ? K : 1];
    for (int i = 0; i < K; ++i) qs[i] = qubits[i];
    double2 acc[D][D];
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
        for (int b = 0; b < D; ++b) acc[a][b] = make_double2(0.0, 0.0);
    for (u64 r = chunk * per + threadIdx.x; r < (chunk + 1) * per; r += blockDim.x) {
        // insert zero bits at the (ascending) qubit positions
        u64 x = r;
        for (int i = 0; i < K; ++i) {
            int q = qs[0];
            if (K == 2) q = (i == 0) ? (qs[0] < qs[1] ? qs[0] : qs[1]) : (qs[0] < qs[1] ? qs[1] : qs[0]);
            const u64 low = (1ull << q) - 1;
            x = ((x & ~low) << 1) | (x & low);
        }
        double2 v[D];
#pragma unroll
        for (int a = 0; a < D; ++a) {
            u64 o = x;
#pragma unroll
            for (int i = 0; i < K; ++i)
                if ((a >> i) & 1) o |= 1ull << qs[i];
            v[a] = st[o];
        }
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
            for (int b = a; b < D; ++b) {       // v[a] * conj(v[b])
                acc[a][b].x += v[a].x * v[b].x + v[a].y * v[b].y;
                acc[a][b].y += v[a].y * v[b].x - v[a].x * v[b].y;
            }
    }
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
        for (int b = a; b < D; ++b) {
            double re = acc[a][b].x, im = acc[a][b].y;
            for (int o = 16; o > 0; o >>= 1) {
                re += __shfl_xor_sync(0xffffffffu, re, o);
                im += __shfl_xor_sync(0xffffffffu, im, o);
            }
            if ((threadIdx.x & 31) == 0) {
                double* dst = (double*)(out + (traj * D + a) * D + b);
                atomicAdd(dst, re);
                if (a != b) {
                    atomicAdd(dst + 1, im);
                    double* dst2 = (double*)(out + (traj * D + b) * D + a);
                    atomicAdd(dst2, re);
                    atomicAdd(dst2 + 1, -im);
                }
            }
        }
}

__global__ void k_sample_rows(const double* __restrict__ probs, long long n_rows, int n_cols, int n_samples,
                              u64 seed, long long traj_offset, int* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows * n_samples) return;
    const long long r = i / n_samples;
    const int s = (int)(i % n_samples);
    const double u = philox_uniform(seed, (uint32_t)s, 1u, (u64)(traj_offset + r));
    const double* p = probs + r * n_cols;
    double cum = 0.0;
    int pick = n_cols - 1;
    for (int c = 0; c < n_cols; ++c) {
        cum += p[c];
        if (cum > u) { pick = c; break; }
    }
    out[i] = pick;
}

__global__ void k_chunk_norms(const double2* __restrict__ state, int n_local, int chunk_bits,
                              double* __restrict__ sums) {
    // one block per (traj, chunk)
    __shared__ double wsum[32];
    const u64 blk = blockIdx.x;
    const double2* st = state + (blk << chunk_bits);
    double acc = 0.0;
    for (u64 x = threadIdx.x; x < (1ull << chunk_bits); x += blockDim.x) {
        const double2 v = st[x];
        acc += v.x * v.x + v.y * v.y;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += wsum[w];
        sums[blk] = s;
    }
}

__global__ void k_sample_states(const double2* __restrict__ state, int n_local, int chunk_bits,
                                const double* __restrict__ sums, long long n_traj, int n_samples, u64 seed,
                                long long traj_offset, const u64* __restrict__ fx, u64* __restrict__ out) {
    // thread per (trajectory, sample): sample s of a trajectory uses philox index s (stream 1)
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_traj * n_samples) return;
    const long long t = i / n_samples;
    const uint32_t smp = (uint32_t)(i % n_samples);
    const u64 nchunks = 1ull << (n_local - chunk_bits);
    const double* s = sums + (u64)t * nchunks;
    double total = 0.0;
    for (u64 c = 0; c < nchunks; ++c) total += s[c];
    const double target = philox_uniform(seed, smp, 1u, (u64)(traj_offset + t)) * total;
    double cum = 0.0;
    u64 c = 0;
    for (; c + 1 < nchunks; ++c) {
        if (cum + s[c] > target) break;
        cum += s[c];
    }
    const double2* st = state + ((u64)t << n_local) + (c << chunk_bits);
    u64 pick = (1ull << chunk_bits) - 1;
    for (u64 x = 0; x < (1ull << chunk_bits); ++x) {
        const double2 v = st[x];
        cum += v.x * v.x + v.y * v.y;
        if (cum > target) { pick = x; break; }
    }
    u64 idx = (c << chunk_bits) | pick;
    if (fx) idx ^= fx[t];
    out[i] = idx;
}

// ---- read-out of a factorised circuit: the small events on a <= 3-qubit density matrix, one thread per trajectory
__global__ void k_readout_small(const __grid_constant__ DtcSmallPlan S, const DtcEvent* __restrict__ ev,
                                const long long* __restrict__ idx, long long n_small, const double2* __restrict__ rdm,
                                const u64* __restrict__ masks, const u64* __restrict__ fx, long long n_traj,
                                double* __restrict__ probs) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_traj) return;
    double pr[DTC_SMALL_DIM];
    small_readout_traj(S, ev, idx, n_small, rdm + (t << (2 * S.n_reg)), masks + t, n_traj, fx[t], pr);
    for (int b = 0; b < (1 << S.m); ++b) probs[(t << S.m) + b] = pr[b];
}

// ---- density matrix (2n-bit vector, index = row + 2^n col)
__global__ void k_rot_cs(double2* __restrict__ v, int nbits, int bit, double c, double s) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1ll << (nbits - 1))) return;
    const u64 lowm = (1ull << bit) - 1;
    const u64 i0 = (((u64)i & ~lowm) << 1) | ((u64)i & lowm);
    const u64 i1 = i0 | (1ull << bit);
    const double2 x0 = v[i0], x1 = v[i1];
    // [[c, -i s], [-i s, c]]
    v[i0] = make_double2(c * x0.x + s * x1.y, c * x0.y - s * x1.x);
    v[i1] = make_double2(c * x1.x + s * x0.y, c * x1.y - s * x0.x);
}

__global__ void k_dm_diag(double2* __restrict__ rho, int n, int n1, const int* __restrict__ q1,
                          const double* __restrict__ a, int n2, const int* __restrict__ qi,
                          const int* __restrict__ qj, const double* __restrict__ b) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1ll << (2 * n))) return;
    const u64 r = (u64)i & ((1ull << n) - 1), c = (u64)i >> n;
    double ang = 0.0;
    for (int k = 0; k < n1; ++k) {
        const int zr = 1 - 2 * (int)((r >> q1[k]) & 1ull), zc = 1 - 2 * (int)((c >> q1[k]) & 1ull);
        ang += 0.5 * a[k] * (double)(zr - zc);
    }
    for (int k = 0; k < n2; ++k) {
        const int zr = 1 - 2 * (int)(((r >> qi[k]) ^ (r >> qj[k])) & 1ull);
        const int zc = 1 - 2 * (int)(((c >> qi[k]) ^ (c >> qj[k])) & 1ull);
        ang += 0.5 * b[k] * (double)(zr - zc);
    }
    double sn, cs;
    sincos(ang, &sn, &cs);
    rho[i] = cmul(rho[i], make_double2(cs, -sn));
}

__global__ void k_dm_channel(double2* __restrict__ rho, int n, int q, double px, double py, double pz) {
    // thread per group of 4 elements spanning (row bit q, col bit q)
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1ll << (2 * n - 2))) return;
    const int b0 = q, b1 = q + n;
    u64 x = (u64)i;
    const u64 low0 = (1ull << b0) - 1;
    x = ((x & ~low0) << 1) | (x & low0);
    const u64 low1 = (1ull << b1) - 1;
    x = ((x & ~low1) << 1) | (x & low1);
    const u64 i00 = x, i11 = x | (1ull << b0) | (1ull << b1), i01 = x | (1ull << b0), i10 = x | (1ull << b1);
    const double dA = 1.0 - px - py, dB = px + py;              // r_q == c_q
    const double oA = 1.0 - px - py - 2.0 * pz, oB = px - py;   // r_q != c_q
    const double2 e00 = rho[i00], e11 = rho[i11], e01 = rho[i01], e10 = rho[i10];
    rho[i00] = make_double2(dA * e00.x + dB * e11.x, dA * e00.y + dB * e11.y);
    rho[i11] = make_double2(dA * e11.x + dB * e00.x, dA * e11.y + dB * e00.y);
    rho[i01] = make_double2(oA * e01.x + oB * e10.x, oA * e01.y + oB * e10.y);
    rho[i10] = make_double2(oA * e10.x + oB * e01.x, oA * e10.y + oB * e01.y);
}


// general single-qubit channel: 4 x 4 complex superoperator on the (row bit q, column bit q) block (csrc/dtc_dm.cuh)
__global__ void k_dm_superop(double2* __restrict__ rho, int n, int q, const __grid_constant__ DmSuperop S) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1ll << (2 * n - 2))) return;
    dm_superop_thread(rho, n, q, S, i);
}

// ---- density matrix, fused: one sweep of rho applies, to up to six qubits at once, the rotation on the row bit, its
// conjugate on the column bit and the Pauli channel that follows (a 2 x 2 block of rho per qubit), with the diagonal layer
// that precedes them folded into the load.  rho is a 2n-bit vector (index = row + 2^n col); a tile holds 2^TB elements:
// optional passive row bits {0,1} (64 B runs) + the row and column bits of the pass's qubits.
// T[x] = exp(-i phi(x)), phi(x) = sum_k a_k z_k(x)/2 + sum_k b_k z_i z_j /2 : rho[r,c] *= T[r] conj(T[c])
__global__ void k_dm_phase_table(double2* __restrict__ T, int n, const __grid_constant__ DmDiagTerms D) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= (1 << n)) return;
    double ang = 0.0;
    for (int k = 0; k < D.n1; ++k) ang += 0.5 * D.a[k] * (double)(1 - 2 * ((x >> D.q1[k]) & 1));
    for (int k = 0; k < D.n2; ++k) ang += 0.5 * D.b[k] * (double)(1 - 2 * (((x >> D.qi[k]) ^ (x >> D.qj[k])) & 1));
    double sn, cs;
    sincos(ang, &sn, &cs);
    T[x] = make_double2(cs, -sn);
}

__global__ void k_dm_apply_table(double2* __restrict__ rho, int n, const double2* __restrict__ T) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1ll << (2 * n))) return;
    const double2 tr = T[(u64)i & ((1ull << n) - 1)], tc = T[(u64)i >> n];
    rho[i] = cmul(rho[i], cmul(tr, make_double2(tc.x, -tc.y)));
}

// THREADS = 256 (tiles of <= 2^12 elements, three CTAs per SM) or 512 (2^13 elements = 128 KB, one CTA per SM)
template <int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 256 ? 3 : 1)
k_dm_tile(double2* __restrict__ rho, const __grid_constant__ DmTilePass P, const double2* __restrict__ T) {
    extern __shared__ __align__(16) double2 dm_tile[];
    __shared__ u64 off_hi[32];
    constexpr int DTC_DM_THREADS_ = THREADS;
    const int tid = threadIdx.x, ne = 1 << P.tile_bits;
    u64 base = 0;
    for (int k = 0; k < P.seg_n; ++k)
        base |= (((u64)blockIdx.x >> P.seg_src[k]) & ((1ull << P.seg_len[k]) - 1)) << P.seg_dst[k];
    // element i = tid + 256 k of the tile: global offset = offset of tid's eight bits (per thread, once) | offset of k's bits
    u64 off_lo = 0;
#pragma unroll
    for (int l = 0; l < 9; ++l)
        if (l < P.tile_bits && ((tid >> l) & 1)) off_lo |= 1ull << P.tb[l];
    constexpr int LB = THREADS == 256 ? 8 : 9;            // element i = tid + THREADS k: tid carries LB bits, k the rest
    if (tid < 32) {
        u64 o = 0;
        for (int l = LB; l < P.tile_bits; ++l)
            if ((tid >> (l - LB)) & 1) o |= 1ull << P.tb[l];
        off_hi[tid] = o;
    }
    __syncthreads();
    const u64 rmask = (1ull << P.n) - 1;
    const int nk = ne > DTC_DM_THREADS_ ? ne / DTC_DM_THREADS_ : 1;
    {
        // all of a thread's loads are issued before the first one is used (a rolled loop would pay one HBM latency per element)
        double2 v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if (k < nk && tid + DTC_DM_THREADS_ * k < ne) v[k] = __ldcs(rho + (base | off_lo | off_hi[k]));
        if (P.has_diag) {
#pragma unroll
            for (int k = 0; k < 16; ++k)
                if (k < nk && tid + DTC_DM_THREADS_ * k < ne) {
                    const u64 g = base | off_lo | off_hi[k];
                    const double2 tr = __ldg(T + (g & rmask)), tc = __ldg(T + (g >> P.n));
                    v[k] = cmul(v[k], cmul(tr, make_double2(tc.x, -tc.y)));
                }
        }
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if (k < nk && tid + DTC_DM_THREADS_ * k < ne) dm_tile[dm_phys(tid + DTC_DM_THREADS_ * k)] = v[k];
    }
    __syncthreads();
    for (int k = 0; k < P.nq; ++k) {
        const DmQubitOp& Q = P.q[k];
        const int lowr = (1 << Q.lr) - 1, lowc = (1 << Q.lc) - 1;
        const double c = Q.c, s = Q.s;
#pragma unroll 4
        for (int gidx = tid; gidx < (ne >> 2); gidx += DTC_DM_THREADS_) {
            int x = ((gidx & ~lowr) << 1) | (gidx & lowr);          // zero at bit lr
            x = ((x & ~lowc) << 1) | (x & lowc);                    // zero at bit lc (lc > lr)
            const int i00 = dm_phys(x), i10 = dm_phys(x | (1 << Q.lr)), i01 = dm_phys(x | (1 << Q.lc)),
                      i11 = dm_phys(x | (1 << Q.lr) | (1 << Q.lc));  // i<row bit><column bit>
            double2 e00 = dm_tile[i00], e10 = dm_tile[i10], e01 = dm_tile[i01], e11 = dm_tile[i11];
            // rows: [[c, -i s], [-i s, c]] on (r = 0, r = 1) for each column bit
            double2 a00 = make_double2(c * e00.x + s * e10.y, c * e00.y - s * e10.x);
            double2 a10 = make_double2(c * e10.x + s * e00.y, c * e10.y - s * e00.x);
            double2 a01 = make_double2(c * e01.x + s * e11.y, c * e01.y - s * e11.x);
            double2 a11 = make_double2(c * e11.x + s * e01.y, c * e11.y - s * e01.x);
            // columns: the conjugate [[c, +i s], [+i s, c]] on (c = 0, c = 1) for each row bit
            e00 = make_double2(c * a00.x - s * a01.y, c * a00.y + s * a01.x);
            e01 = make_double2(c * a01.x - s * a00.y, c * a01.y + s * a00.x);
            e10 = make_double2(c * a10.x - s * a11.y, c * a10.y + s * a11.x);
            e11 = make_double2(c * a11.x - s * a10.y, c * a11.y + s * a10.x);
            // Pauli channel on the (row bit, column bit) block
            dm_tile[i00] = make_double2(Q.dA * e00.x + Q.dB * e11.x, Q.dA * e00.y + Q.dB * e11.y);
            dm_tile[i11] = make_double2(Q.dA * e11.x + Q.dB * e00.x, Q.dA * e11.y + Q.dB * e00.y);
            dm_tile[i10] = make_double2(Q.oA * e10.x + Q.oB * e01.x, Q.oA * e10.y + Q.oB * e01.y);
            dm_tile[i01] = make_double2(Q.oA * e01.x + Q.oB * e10.x, Q.oA * e01.y + Q.oB * e10.y);
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < 16; ++k)
        if (k < nk && tid + DTC_DM_THREADS_ * k < ne) __stcs(rho + (base | off_lo | off_hi[k]), dm_tile[dm_phys(tid + DTC_DM_THREADS_ * k)]);
}

// Register-resident sweep (csrc/dtc_dm.cuh): 16 elements = two qubits' (row, column) bits per thread; round 0 HBM -> registers
// -> shared memory, middle round shared -> shared, last round shared -> registers -> HBM.  THREADS = 2^(tile_bits - 4).
// Two CTAs per SM for the 2^12 tiles (126 registers; at 80 registers for three CTAs the 16 elements spill), one for 2^13.
template <int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 256 ? 2 : 1)
k_dm_reg(double2* __restrict__ rho, const __grid_constant__ DmRegPass P, const double2* __restrict__ T) {
    extern __shared__ __align__(16) double2 dm_tile[];
    const unsigned base = (unsigned)dm_cta_base((u64)blockIdx.x, P.seg_n, P.seg_src, P.seg_len, P.seg_dst);
    constexpr int LB = THREADS == 256 ? 8 : 9;
    for (int r = 0; r < P.n_rounds; ++r) {
        if (r) __syncthreads();
        dmr_round<LB>(r, (int)threadIdx.x, P, rho, T, base, dm_tile);
    }
}

__global__ void k_dm_probs(const double2* __restrict__ rho, int n, int k, const int* __restrict__ qubits,
                           double* __restrict__ out) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= (1ll << n)) return;
    int bin = 0;
    for (int b = 0; b < k; ++b) bin |= (int)(((u64)r >> qubits[b]) & 1ull) << b;
    atomicAdd(&out[bin], rho[(u64)r + ((u64)r << n)].x);
}

// ---- sharded state: gather/scatter for the global<->local qubit exchange
__global__ void k_shard_pack(const double2* __restrict__ state, double2* __restrict__ out, int n_local, int g,
                             const int* __restrict__ lq, int unpack) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1ll << n_local)) return;
    // packed index = (d << (n_local-g)) | j ; state index = j with bits of d inserted at positions lq[]
    const u64 d = (u64)i >> (n_local - g);
    u64 j = (u64)i & ((1ull << (n_local - g)) - 1);
    u64 used = 0;
    for (int b = 0; b < g; ++b) used |= 1ull << lq[b];
    u64 x = 0;
    int src = 0;
    for (int pos = 0; pos < n_local; ++pos) {
        if ((used >> pos) & 1ull) continue;
        if ((j >> src) & 1ull) x |= 1ull << pos;
        ++src;
    }
    for (int b = 0; b < g; ++b)
        if ((d >> b) & 1ull) x |= 1ull << lq[b];
    if (unpack) ((double2*)state)[x] = out[i];
    else out[i] = state[x];
}


// ---- host side of k_tile_stream
typedef CUresult (*dtc_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static dtc_encode_fn tensor_map_encoder() {
    static dtc_encode_fn fn = []() -> dtc_encode_fn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
        return (dtc_encode_fn)p;
    }();
    return fn;
}

static CUtensorMapL2promotion stream_l2_promotion() {
    static const CUtensorMapL2promotion v = []() {
        const char* e = getenv("DTCSIM_L2_PROMOTION");      // tuning: 0 none, 64, 128, 256 bytes
        const int b = e ? atoi(e) : 0;
        return b == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : b == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
             : b == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
    }();
    return v;
}

// tensor map of the tile {0,1} + [g, g+10) over the whole batch: (8 doubles | bits [2,g) | 32 | 32 | everything above)
static int stream_tensor_map(CUtensorMap* tm, void* state, int n_local, int g, int64_t n_traj, int mode) {
    dtc_encode_fn enc = tensor_map_encoder();
    if (!enc) return fail(DTC_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    CUresult r;
    if (mode == 3) {
        // tile {0..6} + [g, g+5): (256 doubles | bits [7,g) | 32 rows | everything above)
        const cuuint64_t dims[4] = {256, 1ull << (g - 7), 32, (cuuint64_t)n_traj << (n_local - g - 5)};
        const cuuint64_t strides[3] = {2048, 16ull << g, 16ull << (g + 5)};
        const cuuint32_t box[4] = {256, 1, 32, 1};
        const cuuint32_t es[4] = {1, 1, 1, 1};
        r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, state, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, stream_l2_promotion(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        const cuuint64_t dims[5] = {8, 1ull << (g - 2), 32, 32, (cuuint64_t)n_traj << (n_local - g - 10)};
        const cuuint64_t strides[4] = {64, 16ull << g, 16ull << (g + 5), 16ull << (g + 10)};
        const cuuint32_t box[5] = {8, 1, 32, 32, 1};
        const cuuint32_t es[5] = {1, 1, 1, 1, 1};
        r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, state, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, stream_l2_promotion(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) return fail(DTC_ERR_CUDA, "cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
    return DTC_OK;
}

static int g_stream_override = -1;      // dtc_set_stream_engine(); -1: environment / default
static int g_stream_ctas = 0;           // dtc_set_stream_ctas(); 0: one persistent CTA per SM
static int g_store_lag = []() { const char* e = getenv("DTCSIM_STORE_LAG"); return e ? atoi(e) : 1; }();   // out-of-place stores: two in flight per SM
static bool stream_enabled() {
    if (g_stream_override >= 0) return g_stream_override != 0;
    static const bool on = []() {
        const char* e = getenv("DTCSIM_STREAM");      // 0: register-fed k_tile_pass everywhere (A/B measurements)
        return !(e && atoi(e) == 0);
    }();
    return on;
}

// ------------------------------------------------------------------------------------ C ABI
extern "C" {

int dtc_version(void) { return 100; }
const char* dtc_last_error(void) { return g_err.c_str(); }

int dtc_device_count(int* count) {
    if (!count) return fail(DTC_ERR_INVALID, "count is NULL");
    CUDA_TRY(cudaGetDeviceCount(count));
    return DTC_OK;
}

int dtc_program_create(int n_qubits, int n_layers, dtc_program** out) {
    if (!out) return fail(DTC_ERR_INVALID, "out is NULL");
    if (n_qubits < 1 || n_qubits > 62) return fail(DTC_ERR_INVALID, "n_qubits must be in [1, 62]");
    if (n_layers < 1 || n_layers > (1 << 20)) return fail(DTC_ERR_INVALID, "n_layers out of range");
    dtc_program* p = new (std::nothrow) dtc_program();
    if (!p) return fail(DTC_ERR_NOMEM, "out of host memory");
    p->h.n_qubits = n_qubits;
    p->h.n_layers = n_layers;
    p->h.n_exec_layers = n_layers;
    *out = p;
    return DTC_OK;
}

int dtc_program_destroy(dtc_program* p) {
    if (!p) return DTC_OK;
    if (p->h.device >= 0) {
        DeviceGuard guard(p->h.device);
        // kernels of this program may still be queued on the caller's stream: the tables are freed in stream order
        // AFTER the last use (no host synchronisation)
        if (p->used && p->last_use && g_dev[p->h.device].ready) cudaStreamWaitEvent(g_dev[p->h.device].upload, p->last_use, 0);
        table_free(p->d_events, p->h.device);
        table_free(p->d_layers, p->h.device);
        table_free(p->d_small_idx, p->h.device);
        table_free(p->d_spasses, p->h.device);
        if (p->ev0) cudaEventDestroy(p->ev0);
        if (p->ev1) cudaEventDestroy(p->ev1);
        for (cudaEvent_t e : p->pass_ev) cudaEventDestroy(e);
        if (p->uploaded) cudaEventDestroy(p->uploaded);
        if (p->last_use) cudaEventDestroy(p->last_use);
    }
    delete p;
    return DTC_OK;
}

int dtc_program_set_events(dtc_program* p, int64_t n, const int32_t* type, const int32_t* layer,
                           const int32_t* q0, const int32_t* q1, const int32_t* slot, const double* val,
                           const double* probs, double global_phase) {
    if (!p) return fail(DTC_ERR_INVALID, "program is NULL");
    if (p->h.finalized) return fail(DTC_ERR_INVALID, "program already finalized");
    std::string err;
    if (!dtc_stage_events(p->h, n, type, layer, q0, q1, slot, val, probs, global_phase, err))
        return fail(DTC_ERR_INVALID, err);
    return DTC_OK;
}

int dtc_program_set_exec_layers(dtc_program* p, int n_exec_layers) {
    if (!p || p->h.finalized) return fail(DTC_ERR_INVALID, "program is NULL or finalized");
    if (n_exec_layers < 1 || n_exec_layers > p->h.n_layers) return fail(DTC_ERR_INVALID, "n_exec_layers out of range");
    p->h.n_exec_layers = n_exec_layers;
    return DTC_OK;
}

int dtc_program_set_readout_hint(dtc_program* p, int bit) {
    if (!p || p->h.finalized) return fail(DTC_ERR_INVALID, "program is NULL or finalized");
    p->h.readout_bit = bit;
    return DTC_OK;
}

int dtc_program_finalize(dtc_program* p, int device, int engine, int n_local) {
    if (!p) return fail(DTC_ERR_INVALID, "program is NULL");
    if (p->h.finalized) return fail(DTC_ERR_INVALID, "program already finalized");
    if (n_local < 1 || n_local > p->h.n_qubits) return fail(DTC_ERR_INVALID, "n_local out of range");
    if (n_local > 40) return fail(DTC_ERR_INVALID, "n_local exceeds 40 qubits");
    p->h.n_local = n_local;
    p->h.device = device;
    std::string err;
    if (!dtc_build_layers(p->h, err)) return fail(DTC_ERR_INVALID, err);
    if (engine == DTC_ENGINE_AUTO) engine = (n_local >= DTC_TILE_BITS) ? DTC_ENGINE_TILE : DTC_ENGINE_GENERIC;
    if (engine == DTC_ENGINE_TILE && n_local < DTC_TILE_BITS)
        return fail(DTC_ERR_INVALID, "tile engine needs n_local >= 12");
    p->h.engine = engine;
    if (engine == DTC_ENGINE_TILE) {
        if (!dtc_schedule_tile_for_readout(p->h, p->h.readout_bit, err)) return fail(DTC_ERR_INVALID, err);
    } else if (engine == DTC_ENGINE_GENERIC) {
        for (int j = 0; j < p->h.n_exec_layers; ++j)
            if (n_local < 64 && (p->h.layers[j].rot_any >> n_local))
                return fail(DTC_ERR_INVALID, "rotation on a non-local qubit");
        dtc_schedule_generic(p->h);
    } else {
        return fail(DTC_ERR_INVALID, "unknown engine");
    }
    DeviceGuard guard(device);
    CUDA_TRY(guard.err);
    const size_t eb = p->h.events.size() * sizeof(DtcEvent), lb = p->h.layers.size() * sizeof(DtcLayer);
    CUDA_TRY(table_upload((void**)&p->d_events, p->h.events.data(), eb, 16, device));
    CUDA_TRY(table_upload((void**)&p->d_layers, p->h.layers.data(), lb, 16, device));
    static_assert(sizeof(DtcStreamPass) % 16 == 0, "builder warps copy pass descriptors in 16 B pieces");
    if (engine == DTC_ENGINE_TILE && !p->h.spasses.empty() && n_local <= 22 && p->h.spasses.size() <= DTC_RESIDENT_MAX_PASSES) {
        bool ok = true;
        for (DtcStreamPass& S : p->h.spasses) {
            if (!S.mode) { ok = false; break; }
            S.tmap_slot = 0;
            if (S.contig) continue;
            int slot = -1;
            for (int i = 0; i < p->n_tmaps; ++i)
                if (p->tmap_mode[i] == S.mode && p->tmap_g[i] == S.g) slot = i;
            if (slot < 0) {
                if (p->n_tmaps == 2) { ok = false; break; }
                slot = p->n_tmaps++;
                p->tmap_mode[slot] = S.mode;
                p->tmap_g[slot] = S.g;
            }
            S.tmap_slot = slot;
        }
        p->resident_ok = ok;
        if (ok)
            CUDA_TRY(table_upload((void**)&p->d_spasses, p->h.spasses.data(), p->h.spasses.size() * sizeof(DtcStreamPass), 16, device));
    }
    CUDA_TRY(cudaEventCreateWithFlags(&p->uploaded, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&p->last_use, cudaEventDisableTiming));
    CUDA_TRY(cudaEventRecord(p->uploaded, g_dev[device].upload));
    static bool attr_set[16] = {false};
    if (device >= 0 && device < 16 && !attr_set[device]) {
        const int smb = (int)sizeof(TileSmem);
        CUDA_TRY(cudaFuncSetAttribute(k_tile_pass<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smb));
        CUDA_TRY(cudaFuncSetAttribute(k_tile_pass<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smb));
        CUDA_TRY(cudaFuncSetAttribute(k_tile_pass<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smb));
        CUDA_TRY(cudaFuncSetAttribute(k_tile_pass<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smb));
        CUDA_TRY(cudaFuncSetAttribute(k_tile_pass<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smb));
        CUDA_TRY(cudaFuncSetAttribute(k_tile_pass<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smb));
        const int ssb = (int)sizeof(StreamSmem) + 128;
        CUDA_TRY(cudaFuncSetAttribute(k_tile_stream<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ssb));
        CUDA_TRY(cudaFuncSetAttribute(k_tile_stream<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ssb));
        CUDA_TRY(cudaFuncSetAttribute(k_tile_stream<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, ssb));
        CUDA_TRY(cudaFuncSetAttribute(k_tile_resident, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ResidentSmem) + 128));
        CUDA_TRY(cudaDeviceGetAttribute(&g_num_sms[device], cudaDevAttrMultiProcessorCount, device));
        attr_set[device] = true;
    }
    p->h.finalized = true;
    return DTC_OK;
}

int dtc_program_num_passes(const dtc_program* p, int* n_passes) {
    if (!p || !n_passes || !p->h.finalized) return fail(DTC_ERR_INVALID, "program not finalized");
    *n_passes = (p->h.engine == DTC_ENGINE_TILE) ? (int)p->h.passes.size() : (int)p->h.gsteps.size();
    return DTC_OK;
}

int dtc_program_workspace_bytes(const dtc_program* p, int64_t n_traj, size_t* bytes) {
    if (!p || !bytes || n_traj < 1) return fail(DTC_ERR_INVALID, "bad argument");
    *bytes = dtc_workspace_bytes(p->h, n_traj);
    return DTC_OK;
}

static void ws_pointers(const DtcProgramHost& h, void* ws, int64_t n_traj, u64** masks, u64** fx, u64** fz, int** ph) {
    u64* base = (u64*)ws;
    *masks = base;
    *fx = base + (size_t)h.n_layers * 4 * n_traj;
    *fz = *fx + n_traj;
    *ph = (int*)(*fz + n_traj);
}

int dtc_program_frames(const dtc_program* p, void* workspace, int64_t n_traj, uint64_t** fx, uint64_t** fz, int32_t** ph) {
    if (!p || !workspace || !fx || !fz || !ph) return fail(DTC_ERR_INVALID, "bad argument");
    u64 *m, *x, *z;
    int* h;
    ws_pointers(p->h, workspace, n_traj, &m, &x, &z, &h);
    *fx = (uint64_t*)x; *fz = (uint64_t*)z; *ph = (int32_t*)h;
    return DTC_OK;
}

// frames of the batch: sign masks of every layer + final frames into the workspace
static int run_frames(dtc_program* p, int64_t n_traj, int64_t traj_offset, uint64_t seed, void* workspace, cudaStream_t s) {
    const DtcProgramHost& h = p->h;
    u64 *masks, *fx, *fz;
    int* ph;
    ws_pointers(h, workspace, n_traj, &masks, &fx, &fz, &ph);
    CUDA_TRY(cudaMemsetAsync(masks, 0, (size_t)h.n_layers * 4 * n_traj * sizeof(u64), s));
    const int fb = 128;
    k_frames<<<(unsigned)((n_traj + fb - 1) / fb), fb, 0, s>>>(p->d_events, (long long)h.events.size(), masks,
                                                              n_traj, traj_offset, seed, fx, fz, ph);
    return DTC_OK;
}

// passes [begin, end) of the tile engine's schedule.  store_last != nullptr: the last pass of the range stores its tiles there
// instead of in place (that pass must run on k_tile_stream with contiguous tiles).  n_ctas > 0 limits the persistent grid.
static int run_tile_passes(dtc_program* p, void* state, void* store_last, int begin, int end, int n_ctas, int64_t n_traj,
                           uint64_t init_index, bool gen_first, bool fused, uint64_t rank_bits, void* workspace, cudaStream_t s) {
    const DtcProgramHost& h = p->h;
    u64 *masks, *fx, *fz;
    int* ph;
    ws_pointers(h, workspace, n_traj, &masks, &fx, &fz, &ph);
    const long long grid = n_traj << (h.n_local - DTC_TILE_BITS);
    if (grid > 0x7fffffffLL) return fail(DTC_ERR_INVALID, "batch too large for one launch");
    static const int pf = []() {
        const char* e = getenv("DTCSIM_PREFETCH_BLOCKS");      // tuning knob; default 148 CTAs ahead (measured best on B200)
        return e ? atoi(e) : 148;
    }();
    const int n_sms = (h.device >= 0 && h.device < 16 && g_num_sms[h.device] > 0) ? g_num_sms[h.device] : 148;
    const bool per_pass = p->profiling && begin == 0 && end == (int)h.passes.size();
    if (per_pass) {
        while ((int)p->pass_ev.size() < end) {
            cudaEvent_t e;
            CUDA_TRY(cudaEventCreate(&e));
            p->pass_ev.push_back(e);
        }
        p->pass_ev_n = end;
    }
    for (int ip = begin; ip < end; ++ip) {
        if (per_pass && ip > 0) CUDA_TRY(cudaEventRecord(p->pass_ev[(size_t)ip - 1], s));
        const DtcTilePass& T = h.passes[(size_t)ip];
        const DtcStreamPass& S = h.spasses[(size_t)ip];
        double2* out = (double2*)state;
        if (store_last && ip + 1 == end) {
            if (!(S.mode && S.contig && stream_enabled()))
                return fail(DTC_ERR_UNSUPPORTED, "out-of-place store needs a streaming pass with contiguous tiles");
            out = (double2*)store_last;
        }
        if (S.mode && stream_enabled()) {
            // TMA-fed streaming engine: one persistent CTA per SM
            alignas(64) CUtensorMap tm;
            memset(&tm, 0, sizeof(tm));
            if (!S.contig) {
                const int rc = stream_tensor_map(&tm, state, h.n_local, S.g, n_traj, S.mode);
                if (rc != DTC_OK) return rc;
            }
            int max_ctas = (g_stream_ctas > 0 && g_stream_ctas < n_sms) ? g_stream_ctas : n_sms;
            if (n_ctas > 0 && n_ctas < max_ctas) max_ctas = n_ctas;
            const unsigned sgrid = (unsigned)(grid < max_ctas ? grid : max_ctas);
            const size_t ssb = sizeof(StreamSmem) + 128;
            const u64 init = (ip == 0 && gen_first) ? (u64)init_index : (u64)DTC_INIT_KEEP;
            const int lag = (out != (double2*)state && g_store_lag && !(ip == 0 && gen_first)) ? 1 : 0;
            double2* rdm_out = nullptr;
            if (fused && ip + 1 == (int)h.passes.size()) {
                rdm_out = (double2*)((char*)workspace + dtc_workspace_rdm_offset(h, n_traj));
                CUDA_TRY(cudaMemsetAsync(rdm_out, 0, sizeof(double2) * 4 * (size_t)n_traj, s));
            }
            if (S.mode == 1)
                k_tile_stream<1><<<sgrid, DTC_STREAM_THREADS, ssb, s>>>((double2*)state, tm, S, p->d_layers, masks, n_traj, rank_bits, grid, init,
                                                                       rdm_out, p->fused_local_bit, out, lag);
            else if (S.mode == 2)
                k_tile_stream<2><<<sgrid, DTC_STREAM_THREADS, ssb, s>>>((double2*)state, tm, S, p->d_layers, masks, n_traj, rank_bits, grid, init,
                                                                       rdm_out, p->fused_local_bit, out, lag);
            else
                k_tile_stream<3><<<sgrid, DTC_STREAM_THREADS, ssb, s>>>((double2*)state, tm, S, p->d_layers, masks, n_traj, rank_bits, grid, init,
                                                                       rdm_out, p->fused_local_bit, out, lag);
            continue;
        }
        const bool hx = T.layerD >= 0 && T.nX > 0;
#define DTC_LAUNCH(S, X)                                                                               \
    k_tile_pass<S, X><<<(unsigned)grid, DTC_THREADS, sizeof(TileSmem), s>>>((double2*)state, T, p->d_layers, masks, \
                                                                            n_traj, rank_bits, pf)
        switch (T.s2_lo * 2 + (hx ? 1 : 0)) {
            case 0: DTC_LAUNCH(0, false); break;
            case 1: DTC_LAUNCH(0, true); break;
            case 2: DTC_LAUNCH(1, false); break;
            case 3: DTC_LAUNCH(1, true); break;
            case 4: DTC_LAUNCH(2, false); break;
            default: DTC_LAUNCH(2, true); break;
        }
#undef DTC_LAUNCH
    }
    return DTC_OK;
}

int dtc_program_run(dtc_program* p, void* state, int64_t n_traj, int64_t traj_offset, uint64_t seed,
                    uint64_t init_index, uint64_t rank_bits, void* workspace, size_t workspace_bytes, void* stream) {
    if (!p || !p->h.finalized) return fail(DTC_ERR_INVALID, "program not finalized");
    if (!state || !workspace || n_traj < 1) return fail(DTC_ERR_INVALID, "bad argument");
    const DtcProgramHost& h = p->h;
    if (workspace_bytes < dtc_workspace_bytes(h, n_traj)) return fail(DTC_ERR_INVALID, "workspace too small");
    const bool keep = init_index == DTC_INIT_KEEP, zero = init_index == DTC_INIT_ZERO;
    if (!keep && !zero && (init_index >> h.n_local)) return fail(DTC_ERR_INVALID, "init_index out of range");
    cudaStream_t s = (cudaStream_t)stream;
    DeviceGuard guard(h.device);
    CUDA_TRY(guard.err);
    CUDA_TRY(cudaStreamWaitEvent(s, p->uploaded, 0));
    int rc = run_frames(p, n_traj, traj_offset, seed, workspace, s);
    if (rc != DTC_OK) return rc;
    u64 *masks, *fx, *fz;
    int* ph;
    ws_pointers(h, workspace, n_traj, &masks, &fx, &fz, &ph);
    const bool fused = p->fuse_rdm && p->fused_local_bit >= 0 && stream_enabled();
    // the first pass can generate the initial state itself (no memset, no read) when it runs on k_tile_stream
    const bool gen_first = !keep && h.engine == DTC_ENGINE_TILE && !h.spasses.empty() && h.spasses[0].mode && stream_enabled();
    if (!keep && !gen_first) {
        const size_t sbytes = ((size_t)n_traj << h.n_local) * sizeof(double2);
        CUDA_TRY(cudaMemsetAsync(state, 0, sbytes, s));
        if (!zero)
            k_init_basis<<<(unsigned)((n_traj + 127) / 128), 128, 0, s>>>((double2*)state, h.n_local, n_traj, init_index);
    }
    p->last_gen_first = gen_first;
    p->last_fused = fused && h.engine == DTC_ENGINE_TILE;
    p->last_resident = false;
    p->last_kernel_launches = 1 + ((h.engine == DTC_ENGINE_TILE) ? (int)h.passes.size() : (int)h.gsteps.size());
    if (p->profiling) CUDA_TRY(cudaEventRecord(p->ev0, s));
    p->last_launches = (h.engine == DTC_ENGINE_TILE) ? (int)h.passes.size() : (int)h.gsteps.size();
    if (h.engine == DTC_ENGINE_TILE) {
        rc = run_tile_passes(p, state, nullptr, 0, (int)h.passes.size(), 0, n_traj, init_index, gen_first, fused, rank_bits, workspace, s);
        if (rc != DTC_OK) return rc;
    } else {
        for (const DtcGenericStep& g : h.gsteps) {
            if (g.kind == 0) {
                const long long np = n_traj << (h.n_local - 1);
                k_generic_rot<<<(unsigned)((np + 255) / 256), 256, 0, s>>>((double2*)state, h.n_local, g.q, g.t,
                                                                         masks + (size_t)(g.layer * 4) * n_traj, n_traj);
            } else {
                const long long ne = n_traj << h.n_local;
                k_generic_diag<<<(unsigned)((ne + 255) / 256), 256, 0, s>>>((double2*)state, h.n_local, p->d_layers + g.layer,
                                                                          masks + (size_t)(g.layer * 4) * n_traj, n_traj, rank_bits);
            }
        }
    }
    if (p->profiling) CUDA_TRY(cudaEventRecord(p->ev1, s));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(p->last_use, s));
    p->used = true;
    return DTC_OK;
}

int dtc_program_prepare(dtc_program* p, int64_t n_traj, int64_t traj_offset, uint64_t seed, void* workspace,
                        size_t workspace_bytes, void* stream) {
    if (!p || !p->h.finalized) return fail(DTC_ERR_INVALID, "program not finalized");
    if (!workspace || n_traj < 1) return fail(DTC_ERR_INVALID, "bad argument");
    if (workspace_bytes < dtc_workspace_bytes(p->h, n_traj)) return fail(DTC_ERR_INVALID, "workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    DeviceGuard guard(p->h.device);
    CUDA_TRY(guard.err);
    CUDA_TRY(cudaStreamWaitEvent(s, p->uploaded, 0));
    const int rc = run_frames(p, n_traj, traj_offset, seed, workspace, s);
    if (rc != DTC_OK) return rc;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(p->last_use, s));
    p->used = true;
    return DTC_OK;
}

int dtc_program_run_passes(dtc_program* p, void* state, void* store_last, int pass_begin, int pass_end, int n_ctas,
                           int64_t n_traj, uint64_t init_index, uint64_t rank_bits, void* workspace, size_t workspace_bytes,
                           void* stream) {
    if (!p || !p->h.finalized) return fail(DTC_ERR_INVALID, "program not finalized");
    if (!state || !workspace || n_traj < 1) return fail(DTC_ERR_INVALID, "bad argument");
    const DtcProgramHost& h = p->h;
    if (h.engine != DTC_ENGINE_TILE) return fail(DTC_ERR_UNSUPPORTED, "pass ranges exist for the tile engine only");
    if (pass_begin < 0 || pass_end > (int)h.passes.size() || pass_begin >= pass_end) return fail(DTC_ERR_INVALID, "bad pass range");
    if (workspace_bytes < dtc_workspace_bytes(h, n_traj)) return fail(DTC_ERR_INVALID, "workspace too small");
    const bool keep = init_index == DTC_INIT_KEEP, zero = init_index == DTC_INIT_ZERO;
    if (!keep && !zero && (init_index >> h.n_local)) return fail(DTC_ERR_INVALID, "init_index out of range");
    if (!keep && pass_begin != 0) return fail(DTC_ERR_INVALID, "only the first pass can start from a basis state");
    cudaStream_t s = (cudaStream_t)stream;
    DeviceGuard guard(h.device);
    CUDA_TRY(guard.err);
    CUDA_TRY(cudaStreamWaitEvent(s, p->uploaded, 0));
    const bool gen_first = !keep && h.spasses[0].mode && stream_enabled();
    if (!keep && !gen_first) {
        CUDA_TRY(cudaMemsetAsync(state, 0, ((size_t)n_traj << h.n_local) * sizeof(double2), s));
        if (!zero)
            k_init_basis<<<(unsigned)((n_traj + 127) / 128), 128, 0, s>>>((double2*)state, h.n_local, n_traj, init_index);
    }
    const int rc = run_tile_passes(p, state, store_last, pass_begin, pass_end, n_ctas, n_traj, init_index, gen_first, false, rank_bits,
                                   workspace, s);
    if (rc != DTC_OK) return rc;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(p->last_use, s));
    p->used = true;
    return DTC_OK;
}

int dtc_program_pass_info(const dtc_program* p, int pass, int* streaming, int* contiguous) {
    if (!p || !p->h.finalized || !streaming || !contiguous) return fail(DTC_ERR_INVALID, "bad argument");
    if (p->h.engine != DTC_ENGINE_TILE || pass < 0 || pass >= (int)p->h.spasses.size()) return fail(DTC_ERR_INVALID, "no such pass");
    *streaming = p->h.spasses[(size_t)pass].mode != 0 && stream_enabled();
    *contiguous = p->h.spasses[(size_t)pass].contig;
    return DTC_OK;
}


// ---- resident execution (k_tile_resident): read-out-only runs of programs whose passes all stream
static size_t g_resident_bytes = 64ull << 20;      // dtc_set_resident_bytes(): state bytes kept in flight (L2 is 126 MB)

static int resident_group(const dtc_program* p, int64_t n_traj) {
    const size_t sb = sizeof(double2) << p->h.n_local;
    long long G = (long long)(g_resident_bytes / sb);
    if (G < 1) G = 1;
    if (G > 64) G = 64;
    if (G > n_traj) G = n_traj;
    return (int)G;
}

int dtc_set_resident_bytes(size_t bytes) {
    if (bytes < (1u << 20)) return fail(DTC_ERR_INVALID, "resident bytes must be at least 1 MiB");
    g_resident_bytes = bytes;
    return DTC_OK;
}

int dtc_program_resident_info(const dtc_program* p, int64_t n_traj, int* eligible, int* group, size_t* scratch_bytes) {
    if (!p || !p->h.finalized || n_traj < 1 || !eligible || !group || !scratch_bytes) return fail(DTC_ERR_INVALID, "bad argument");
    const int G = resident_group(p, n_traj);
    const long long n_groups = (n_traj + G - 1) / G;
    const bool small = n_groups * (long long)p->h.passes.size() * G < (1ll << 30);
    *eligible = (p->resident_ok && p->fuse_rdm && p->fused_local_bit >= 0 && stream_enabled() && small) ? 1 : 0;
    *group = G;
    *scratch_bytes = ((size_t)G << p->h.n_local) * sizeof(double2);
    return DTC_OK;
}

int dtc_program_run_resident(dtc_program* p, void* scratch, size_t scratch_bytes, int64_t n_traj, int64_t traj_offset,
                             uint64_t seed, uint64_t init_index, uint64_t rank_bits, void* workspace, size_t workspace_bytes,
                             void* stream) {
    if (!p || !p->h.finalized) return fail(DTC_ERR_INVALID, "program not finalized");
    if (!scratch || !workspace || n_traj < 1) return fail(DTC_ERR_INVALID, "bad argument");
    int eligible = 0, G = 0;
    size_t need = 0;
    if (dtc_program_resident_info(p, n_traj, &eligible, &G, &need) != DTC_OK) return DTC_ERR_INVALID;
    if (!eligible) return fail(DTC_ERR_UNSUPPORTED, "program is not eligible for resident execution (needs the fused read-out and streaming passes only)");
    if (scratch_bytes < need) return fail(DTC_ERR_INVALID, "scratch buffer too small");
    const DtcProgramHost& h = p->h;
    if (workspace_bytes < dtc_workspace_bytes(h, n_traj)) return fail(DTC_ERR_INVALID, "workspace too small");
    if (init_index == DTC_INIT_KEEP) return fail(DTC_ERR_UNSUPPORTED, "resident execution starts from a basis state");
    if (init_index != DTC_INIT_ZERO && (init_index >> h.n_local)) return fail(DTC_ERR_INVALID, "init_index out of range");
    cudaStream_t s = (cudaStream_t)stream;
    DeviceGuard guard(h.device);
    CUDA_TRY(guard.err);
    CUDA_TRY(cudaStreamWaitEvent(s, p->uploaded, 0));
    u64 *masks, *fx, *fz;
    int* ph;
    ws_pointers(h, workspace, n_traj, &masks, &fx, &fz, &ph);
    CUDA_TRY(cudaMemsetAsync(masks, 0, (size_t)h.n_layers * 4 * n_traj * sizeof(u64), s));
    const int fb = 128;
    k_frames<<<(unsigned)((n_traj + fb - 1) / fb), fb, 0, s>>>(p->d_events, (long long)h.events.size(), masks,
                                                              n_traj, traj_offset, seed, fx, fz, ph);
    ResidentPlan R;
    memset(&R, 0, sizeof(R));
    R.n_local = h.n_local;
    R.n_passes = (int)h.passes.size();
    R.G = G;
    R.n_traj = n_traj;
    R.n_groups = (n_traj + G - 1) / G;
    R.last_G = (int)(n_traj - (R.n_groups - 1) * G);
    R.nt_bits = h.n_local - DTC_TILE_BITS;
    R.gen_first = 1;
    R.fused_last = 1;
    R.rdm_local_bit = p->fused_local_bit;
    R.items_per_group = (long long)R.n_passes * G << R.nt_bits;
    R.total_items = (R.n_groups - 1) * R.items_per_group + ((long long)R.n_passes * R.last_G << R.nt_bits);
    R.init_index = init_index;
    R.rank_bits = rank_bits;
    double2* rdm_out = (double2*)((char*)workspace + dtc_workspace_rdm_offset(h, n_traj));
    int* cnt = (int*)((char*)workspace + dtc_workspace_cnt_offset(h, n_traj));
    CUDA_TRY(cudaMemsetAsync(rdm_out, 0, sizeof(double2) * 4 * (size_t)n_traj, s));
    CUDA_TRY(cudaMemsetAsync(cnt, 0, sizeof(int) * (size_t)(R.n_groups * R.n_passes * G), s));
    alignas(64) CUtensorMap tm[2];
    memset(tm, 0, sizeof(tm));
    for (int i = 0; i < p->n_tmaps; ++i) {
        const int rc = stream_tensor_map(&tm[i], scratch, h.n_local, p->tmap_g[i], G, p->tmap_mode[i]);
        if (rc != DTC_OK) return rc;
    }
    const int n_sms = (h.device >= 0 && h.device < 16 && g_num_sms[h.device] > 0) ? g_num_sms[h.device] : 148;
    const int max_ctas = (g_stream_ctas > 0 && g_stream_ctas < n_sms) ? g_stream_ctas : n_sms;
    const unsigned grid = (unsigned)(R.total_items < max_ctas ? R.total_items : max_ctas);
    if (p->profiling) CUDA_TRY(cudaEventRecord(p->ev0, s));
    double2* st = (double2*)scratch;
    const DtcStreamPass* dsp = p->d_spasses;
    const DtcLayer* dl = p->d_layers;
    const u64* dm = masks;
    static thread_local ResidentPasses RP;
    for (size_t i = 0; i < h.spasses.size(); ++i) {
        const DtcStreamPass& S = h.spasses[i];
        DtcResidentPass& Q = RP.p[i];
        Q.mode = (unsigned char)S.mode; Q.contig = (unsigned char)S.contig; Q.g = (unsigned char)S.g;
        Q.tmap_slot = (unsigned char)S.tmap_slot;
        Q.layerA = (short)S.layerA; Q.layerD = (short)S.layerD; Q.layerB = (short)S.layerB; Q.n_local = (short)S.n_local;
        for (int l = 0; l < DTC_TILE_BITS; ++l) { Q.tb[l] = (unsigned char)S.tb[l]; Q.t1[l] = S.t1[l]; Q.t2[l] = S.t2[l]; }
        Q.tile_mask = S.tile_mask;
    }
    void* args[] = {&st, &tm[0], &tm[1], &R, &RP, &dsp, &dl, &dm, &cnt, &rdm_out};
    // cooperative launch: all CTAs are co-resident by construction (they wait on one another's progress)
    CUDA_TRY(cudaLaunchCooperativeKernel((const void*)k_tile_resident, dim3(grid), dim3(DTC_STREAM_THREADS), args,
                                         sizeof(ResidentSmem) + 128, s));
    if (p->profiling) CUDA_TRY(cudaEventRecord(p->ev1, s));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(p->last_use, s));
    p->used = true;
    p->last_gen_first = true;
    p->last_fused = true;
    p->last_resident = true;
    p->last_launches = R.n_passes;
    p->last_kernel_launches = 2;
    return DTC_OK;
}

int dtc_program_last_run_info(const dtc_program* p, int* resident, int* kernel_launches) {
    if (!p || !resident || !kernel_launches) return fail(DTC_ERR_INVALID, "bad argument");
    *resident = p->last_resident ? 1 : 0;
    *kernel_launches = p->last_kernel_launches;
    return DTC_OK;
}

#ifdef DTC_STREAM_TIMING
/* tuning builds only: out[16] = cycles summed over CTAs -- TMA driver {wait done, wait store read, issue, drain},
 * table builder {build1+2, wait done, build3, -}, compute warpgroup 0 and 1 {wait full, phase 1, phase 2, phase 3}; resets. */
int dtc_debug_stream_timing(unsigned long long* out) {
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpyFromSymbol(out, g_stream_prof, sizeof(unsigned long long) * 16));
    unsigned long long z[16] = {0};
    CUDA_TRY(cudaMemcpyToSymbol(g_stream_prof, z, sizeof(z)));
    return DTC_OK;
}
#endif

int dtc_program_set_readout(dtc_program* p, int64_t n_small, const int64_t* small_events, int n_reg, const int32_t* reg_bits,
                            int n_elim, const int32_t* elim_bits, int m, const int32_t* measure_bits) {
    if (!p || !p->h.finalized) return fail(DTC_ERR_INVALID, "program not finalized");
    if (n_small < 0 || (n_small > 0 && !small_events) || n_reg < 0 || n_reg > 2 || n_elim < 1 || n_reg + n_elim > DTC_SMALL_MAXQ ||
        m < 1 || m > DTC_SMALL_MAXQ || (n_reg > 0 && !reg_bits) || !elim_bits || !measure_bits)
        return fail(DTC_ERR_INVALID, "read-out plan: at most 3 qubits (<= 2 register bits), 1..3 measured bits");
    DtcSmallPlan S;
    memset(&S, 0, sizeof(S));
    S.nq = n_reg + n_elim; S.n_reg = n_reg; S.m = m;
    for (int i = 0; i < n_reg; ++i) S.bits[i] = reg_bits[i];
    for (int i = 0; i < n_elim; ++i) S.bits[n_reg + i] = elim_bits[i];
    for (int i = 0; i < m; ++i) {
        S.meas_bit[i] = measure_bits[i];
        S.meas_pos[i] = small_pos(S, measure_bits[i]);
        if (S.meas_pos[i] < 0) return fail(DTC_ERR_INVALID, "read-out plan: measured bit outside the small register");
    }
    for (int64_t e = 0; e < n_small; ++e)
        if (small_events[e] < 0 || small_events[e] >= (int64_t)p->h.events.size())
            return fail(DTC_ERR_INVALID, "read-out plan: event index out of range");
    DeviceGuard guard(p->h.device);
    CUDA_TRY(guard.err);
    if (p->used) CUDA_TRY(cudaStreamWaitEvent(g_dev[p->h.device].upload, p->last_use, 0));
    table_free(p->d_small_idx, p->h.device);
    p->d_small_idx = nullptr;
    CUDA_TRY(table_upload((void**)&p->d_small_idx, small_events, sizeof(long long) * (size_t)n_small, 16, p->h.device));
    CUDA_TRY(cudaEventRecord(p->uploaded, g_dev[p->h.device].upload));
    p->small = S;
    p->n_small = n_small;
    return DTC_OK;
}

int dtc_program_set_fused_rdm(dtc_program* p, int enable, int* active) {
    if (!p || !p->h.finalized) return fail(DTC_ERR_INVALID, "program not finalized");
    p->fuse_rdm = false;
    p->fused_local_bit = -1;
    if (enable && p->n_small >= 0 && p->small.n_reg == 1 && p->h.engine == DTC_ENGINE_TILE && !p->h.spasses.empty()) {
        const DtcStreamPass& S = p->h.spasses.back();
        if (S.mode)
            for (int l = 0; l < DTC_TILE_BITS; ++l)
                if (S.tb[l] == p->small.bits[0]) p->fused_local_bit = l;
        p->fuse_rdm = p->fused_local_bit >= 0;
    }
    if (active) *active = p->fuse_rdm && stream_enabled();
    return DTC_OK;
}

int dtc_program_fused_rdm(const dtc_program* p, void* workspace, int64_t n_traj, void** rdm) {
    if (!p || !workspace || !rdm) return fail(DTC_ERR_INVALID, "bad argument");
    *rdm = (p->fuse_rdm && stream_enabled()) ? (void*)((char*)workspace + dtc_workspace_rdm_offset(p->h, n_traj)) : nullptr;
    return DTC_OK;
}

int dtc_program_readout(const dtc_program* p, const void* rdm, void* workspace, int64_t n_traj, double* probs, void* stream) {
    if (!p || p->n_small < 0) return fail(DTC_ERR_INVALID, "no read-out plan set (dtc_program_set_readout)");
    if (!rdm || !workspace || !probs || n_traj < 1) return fail(DTC_ERR_INVALID, "bad argument");
    u64 *masks, *fx, *fz;
    int* ph;
    ws_pointers(p->h, workspace, n_traj, &masks, &fx, &fz, &ph);
    DeviceGuard guard(p->h.device);
    CUDA_TRY(guard.err);
    CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, p->uploaded, 0));
    k_readout_small<<<(unsigned)((n_traj + 63) / 64), 64, 0, (cudaStream_t)stream>>>(p->small, p->d_events, p->d_small_idx, p->n_small,
                                                                                    (const double2*)rdm, masks, fx, n_traj, probs);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(p->last_use, (cudaStream_t)stream));
    const_cast<dtc_program*>(p)->used = true;
    return DTC_OK;
}

int dtc_set_high_stride_bit(int bit) {
    if (bit < 7 || bit > 62) return fail(DTC_ERR_INVALID, "high-stride bit must be in [7, 62]");
    g_dtc_high_stride_bit = bit;
    return DTC_OK;
}

int dtc_set_stream_engine(int enable) {
    g_stream_override = enable < 0 ? -1 : (enable != 0);
    return DTC_OK;
}

int dtc_set_stream_ctas(int n_ctas) {
    if (n_ctas < 0) return fail(DTC_ERR_INVALID, "n_ctas must be >= 0");
    g_stream_ctas = n_ctas;
    return DTC_OK;
}

int dtc_program_num_stream_passes(const dtc_program* p, int* n_passes) {
    if (!p || !n_passes || !p->h.finalized) return fail(DTC_ERR_INVALID, "program not finalized");
    int n = 0;
    for (const DtcStreamPass& S : p->h.spasses) n += S.mode != 0;
    *n_passes = n;
    return DTC_OK;
}

int dtc_program_last_run_flags(const dtc_program* p, int* gen_first, int* fused_rdm) {
    if (!p || !gen_first || !fused_rdm) return fail(DTC_ERR_INVALID, "bad argument");
    *gen_first = p->last_gen_first ? 1 : 0;
    *fused_rdm = p->last_fused ? 1 : 0;
    return DTC_OK;
}

int dtc_program_set_profiling(dtc_program* p, int enable) {
    if (!p) return fail(DTC_ERR_INVALID, "program is NULL");
    if (enable && !p->ev0) {
        CUDA_TRY(cudaEventCreate(&p->ev0));
        CUDA_TRY(cudaEventCreate(&p->ev1));
    }
    p->profiling = enable != 0;
    return DTC_OK;
}

int dtc_program_pass_times(dtc_program* p, float* ms, int* modes, int cap, int* n_passes) {
    if (!p || !ms || !modes || !n_passes || !p->ev1) return fail(DTC_ERR_INVALID, "profiling not enabled");
    CUDA_TRY(cudaEventSynchronize(p->ev1));
    const int n = p->pass_ev_n;
    *n_passes = n;
    for (int i = 0; i < n && i < cap; ++i) {
        cudaEvent_t a = i == 0 ? p->ev0 : p->pass_ev[(size_t)i - 1];
        cudaEvent_t b = i + 1 == n ? p->ev1 : p->pass_ev[(size_t)i];
        CUDA_TRY(cudaEventElapsedTime(&ms[i], a, b));
        modes[i] = p->h.spasses[(size_t)i].mode;
    }
    return DTC_OK;
}

int dtc_program_pass_time(dtc_program* p, float* ms, int* n_launches) {
    if (!p || !ms || !n_launches || !p->ev1) return fail(DTC_ERR_INVALID, "profiling not enabled");
    CUDA_TRY(cudaEventSynchronize(p->ev1));
    CUDA_TRY(cudaEventElapsedTime(ms, p->ev0, p->ev1));
    *n_launches = p->last_launches;
    return DTC_OK;
}

int dtc_materialize(void* state, int n_local, int64_t n_traj, const uint64_t* fx, const uint64_t* fz,
                    const int32_t* ph, void* scratch, void* stream) {
    (void)scratch;
    if (!state || !fx || !fz || !ph) return fail(DTC_ERR_INVALID, "bad argument");
    DeviceGuard guard(device_of(state));
    CUDA_TRY(guard.err);
    const long long ne = n_traj << n_local;
    k_materialize<<<(unsigned)((ne + 255) / 256), 256, 0, (cudaStream_t)stream>>>((double2*)state, n_local, n_traj,
                                                                                 (const u64*)fx, (const u64*)fz, ph);
    CUDA_TRY(cudaGetLastError());
    return DTC_OK;
}

static int chunks_for(int n_local) {
    int cb = n_local - 14;            // 2^14 amplitudes per block
    if (cb < 0) cb = 0;
    if (cb > 10) cb = 10;
    return 1 << cb;
}

int dtc_probs(const void* state, int n_local, int64_t n_traj, int k, const int32_t* qubits,
              const uint64_t* fx, double* out, void* stream) {
    if (!state || !out || k < 0 || k > 12 || (k > 0 && !qubits)) return fail(DTC_ERR_INVALID, "bad argument (k <= 12)");
    DeviceGuard guard(device_of(state));
    CUDA_TRY(guard.err);
    cudaStream_t s = (cudaStream_t)stream;
    int* dq = nullptr;
    CUDA_TRY(cudaMallocAsync(&dq, sizeof(int) * (k ? k : 1), s));
    if (k) CUDA_TRY(cudaMemcpyAsync(dq, qubits, sizeof(int) * k, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(double) * ((size_t)n_traj << k), s));
    const int cpt = chunks_for(n_local);
    k_probs<<<(unsigned)(n_traj * cpt), 256, sizeof(double) << k, s>>>((const double2*)state, n_local, n_traj, k, dq,
                                                                     (const u64*)fx, out, cpt);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaFreeAsync(dq, s));
    return DTC_OK;
}

int dtc_rdm(const void* state, int n_local, int64_t n_traj, int k, const int32_t* qubits, void* out, void* stream) {
    if (!state || !out || k < 0 || k > 2 || k > n_local || (k > 0 && !qubits)) return fail(DTC_ERR_INVALID, "bad argument (k <= 2)");
    DeviceGuard guard(device_of(state));
    CUDA_TRY(guard.err);
    if (k == 2 && qubits[0] == qubits[1]) return fail(DTC_ERR_INVALID, "duplicate qubit");
    cudaStream_t s = (cudaStream_t)stream;
    int* dq = nullptr;
    CUDA_TRY(cudaMallocAsync(&dq, sizeof(int) * 2, s));
    if (k) CUDA_TRY(cudaMemcpyAsync(dq, qubits, sizeof(int) * k, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(double2) * ((size_t)n_traj << (2 * k)), s));
    int cb = n_local - k - 13;
    if (cb < 0) cb = 0;
    if (cb > 8) cb = 8;
    const int cpt = 1 << cb;
    const unsigned grid = (unsigned)(n_traj * cpt);
    if (k == 0) k_rdm<0><<<grid, 256, 0, s>>>((const double2*)state, n_local, n_traj, dq, (double2*)out, cpt);
    else if (k == 1) k_rdm<1><<<grid, 256, 0, s>>>((const double2*)state, n_local, n_traj, dq, (double2*)out, cpt);
    else k_rdm<2><<<grid, 256, 0, s>>>((const double2*)state, n_local, n_traj, dq, (double2*)out, cpt);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaFreeAsync(dq, s));
    return DTC_OK;
}

int dtc_expect_z(const void* state, int n_local, int64_t n_traj, const uint64_t* fx, double* out, void* stream) {
    if (!state || !out || n_local > DTC_MAXQ) return fail(DTC_ERR_INVALID, "bad argument");
    DeviceGuard guard(device_of(state));
    CUDA_TRY(guard.err);
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(double) * (size_t)n_traj * n_local, s));
    const int cpt = chunks_for(n_local);
    k_expect_z<<<(unsigned)(n_traj * cpt), 128, 0, s>>>((const double2*)state, n_local, n_traj, (const u64*)fx, out, cpt);
    CUDA_TRY(cudaGetLastError());
    return DTC_OK;
}

int dtc_sample_rows(const double* probs, int64_t n_rows, int n_cols, int n_samples, uint64_t seed,
                    int64_t traj_offset, int32_t* out, void* stream) {
    if (!probs || !out || n_rows < 1 || n_cols < 1 || n_samples < 1) return fail(DTC_ERR_INVALID, "bad argument");
    DeviceGuard guard(device_of(probs));
    CUDA_TRY(guard.err);
    const long long n = n_rows * n_samples;
    k_sample_rows<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(probs, n_rows, n_cols, n_samples, seed,
                                                                                traj_offset, out);
    CUDA_TRY(cudaGetLastError());
    return DTC_OK;
}

int dtc_sample_states_multi(const void* state, int n_local, int64_t n_traj, int n_samples, uint64_t seed, int64_t traj_offset,
                            const uint64_t* fx, double* scratch, uint64_t* out, void* stream) {
    if (!state || !scratch || !out || n_traj < 1 || n_samples < 1) return fail(DTC_ERR_INVALID, "bad argument");
    DeviceGuard guard(device_of(state));
    CUDA_TRY(guard.err);
    cudaStream_t s = (cudaStream_t)stream;
    const int cb = n_local < 12 ? n_local : 12;
    const long long nblk = n_traj << (n_local - cb);
    const long long n = n_traj * n_samples;
    k_chunk_norms<<<(unsigned)nblk, 128, 0, s>>>((const double2*)state, n_local, cb, scratch);
    k_sample_states<<<(unsigned)((n + 63) / 64), 64, 0, s>>>((const double2*)state, n_local, cb, scratch, n_traj, n_samples, seed,
                                                            traj_offset, (const u64*)fx, (u64*)out);
    CUDA_TRY(cudaGetLastError());
    return DTC_OK;
}

int dtc_sample_states(const void* state, int n_local, int64_t n_traj, uint64_t seed, int64_t traj_offset,
                      const uint64_t* fx, double* scratch, uint64_t* out, void* stream) {
    return dtc_sample_states_multi(state, n_local, n_traj, 1, seed, traj_offset, fx, scratch, out, stream);
}

// ---- density matrix
int dtc_dm_init(void* rho, int n, uint64_t basis_index, void* stream) {
    if (!rho || n < 1 || n > 13 || (basis_index >> n)) return fail(DTC_ERR_INVALID, "bad argument (n <= 13)");
    DeviceGuard guard(device_of(rho));
    CUDA_TRY(guard.err);
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(rho, 0, sizeof(double2) << (2 * n), s));
    k_init_basis<<<1, 32, 0, s>>>((double2*)rho, 2 * n, 1, basis_index | (basis_index << n));
    CUDA_TRY(cudaGetLastError());
    return DTC_OK;
}

int dtc_dm_rot(void* rho, int n, int qubit, double theta, void* stream) {
    if (!rho || n < 1 || n > 13 || qubit < 0 || qubit >= n) return fail(DTC_ERR_INVALID, "bad argument");
    DeviceGuard guard(device_of(rho));
    CUDA_TRY(guard.err);
    cudaStream_t s = (cudaStream_t)stream;
    const double c = cos(0.5 * theta), sn = sin(0.5 * theta);
    const long long np = 1ll << (2 * n - 1);
    k_rot_cs<<<(unsigned)((np + 255) / 256), 256, 0, s>>>((double2*)rho, 2 * n, qubit, c, sn);       // rows: RX(theta)
    k_rot_cs<<<(unsigned)((np + 255) / 256), 256, 0, s>>>((double2*)rho, 2 * n, qubit + n, c, -sn);  // cols: conj
    CUDA_TRY(cudaGetLastError());
    return DTC_OK;
}

int dtc_dm_diag(void* rho, int n, int n1, const int32_t* q1, const double* a, int n2, const int32_t* qi,
                const int32_t* qj, const double* b, void* stream) {
    if (!rho || n < 1 || n > 13 || n1 < 0 || n2 < 0) return fail(DTC_ERR_INVALID, "bad argument");
    DeviceGuard guard(device_of(rho));
    CUDA_TRY(guard.err);
    cudaStream_t s = (cudaStream_t)stream;
    int *dq1 = nullptr, *dqi = nullptr, *dqj = nullptr;
    double *da = nullptr, *db = nullptr;
    CUDA_TRY(cudaMallocAsync(&dq1, sizeof(int) * (n1 + 1), s));
    CUDA_TRY(cudaMallocAsync(&da, sizeof(double) * (n1 + 1), s));
    CUDA_TRY(cudaMallocAsync(&dqi, sizeof(int) * (n2 + 1), s));
    CUDA_TRY(cudaMallocAsync(&dqj, sizeof(int) * (n2 + 1), s));
    CUDA_TRY(cudaMallocAsync(&db, sizeof(double) * (n2 + 1), s));
    if (n1) {
        CUDA_TRY(cudaMemcpyAsync(dq1, q1, sizeof(int) * n1, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(da, a, sizeof(double) * n1, cudaMemcpyHostToDevice, s));
    }
    if (n2) {
        CUDA_TRY(cudaMemcpyAsync(dqi, qi, sizeof(int) * n2, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(dqj, qj, sizeof(int) * n2, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(db, b, sizeof(double) * n2, cudaMemcpyHostToDevice, s));
    }
    const long long ne = 1ll << (2 * n);
    k_dm_diag<<<(unsigned)((ne + 255) / 256), 256, 0, s>>>((double2*)rho, n, n1, dq1, da, n2, dqi, dqj, db);
    CUDA_TRY(cudaGetLastError());
    // host arrays were pageable: make sure the copies are done before the caller reuses them
    CUDA_TRY(cudaStreamSynchronize(s));
    cudaFreeAsync(dq1, s); cudaFreeAsync(da, s); cudaFreeAsync(dqi, s); cudaFreeAsync(dqj, s); cudaFreeAsync(db, s);
    return DTC_OK;
}

int dtc_dm_pauli_channel(void* rho, int n, int qubit, double px, double py, double pz, void* stream) {
    if (!rho || n < 1 || n > 13 || qubit < 0 || qubit >= n) return fail(DTC_ERR_INVALID, "bad argument");
    DeviceGuard guard(device_of(rho));
    CUDA_TRY(guard.err);
    const long long ng = 1ll << (2 * n - 2);
    k_dm_channel<<<(unsigned)((ng + 255) / 256), 256, 0, (cudaStream_t)stream>>>((double2*)rho, n, qubit, px, py, pz);
    CUDA_TRY(cudaGetLastError());
    return DTC_OK;
}


int dtc_dm_superop(void* rho, int n, int qubit, const double* superop, void* stream) {
    if (!rho || !superop || n < 1 || n > 13 || qubit < 0 || qubit >= n) return fail(DTC_ERR_INVALID, "bad argument");
    DeviceGuard guard(device_of(rho));
    CUDA_TRY(guard.err);
    DmSuperop S;
    for (int k = 0; k < 16; ++k) {
        S.re[k] = superop[2 * k];
        S.im[k] = superop[2 * k + 1];
    }
    const long long ng = 1ll << (2 * n - 2);
    k_dm_superop<<<(unsigned)((ng + 255) / 256), 256, 0, (cudaStream_t)stream>>>((double2*)rho, n, qubit, S);
    CUDA_TRY(cudaGetLastError());
    return DTC_OK;
}

// Whole density-matrix program in one call (replaces a per-gate launch loop on the host side): segments of rotations (type 0),
// diagonal terms (1) and Pauli channels (2) in circuit order.  Rotation + channel segments that follow each other are fused
// per qubit; their qubits are swept in groups of up to six per pass; a diagonal segment becomes a phase table that the next
// pass applies while loading (or one elementwise pass if nothing follows).
int dtc_dm_run(void* rho, int n, int n_seg, const int32_t* seg_type, const int32_t* seg_off, const int32_t* q0,
               const int32_t* q1, const double* val, const double* probs, int* n_sweeps, void* stream) {
    if (!rho || n < 1 || n > 13 || n_seg < 0 || (n_seg > 0 && (!seg_type || !seg_off || !q0 || !q1 || !val || !probs)))
        return fail(DTC_ERR_INVALID, "bad argument (n <= 13)");
    DeviceGuard guard(device_of(rho));
    CUDA_TRY(guard.err);
    cudaStream_t s = (cudaStream_t)stream;
    static bool attr_set[DTC_MAX_DEVICES] = {false};
    {
        int dev = 0;
        CUDA_TRY(cudaGetDevice(&dev));
        if (dev >= 0 && dev < DTC_MAX_DEVICES && !attr_set[dev]) {
            CUDA_TRY(cudaFuncSetAttribute(k_dm_tile<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 << 12));
            CUDA_TRY(cudaFuncSetAttribute(k_dm_tile<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 << 13));
            CUDA_TRY(cudaFuncSetAttribute(k_dm_reg<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 << 12));
            CUDA_TRY(cudaFuncSetAttribute(k_dm_reg<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 << 13));
            attr_set[dev] = true;
        }
    }
    // DTCSIM_DM_REG=0: element-per-thread tiles (k_dm_tile) everywhere; DTCSIM_DM_TILE13=0: 2^12 tiles without a passive bit
    // for the groups that do not hold qubit 0 (16 B runs)
    static const DmPlanOptions opt = []() {
        DmPlanOptions o;
        const char* e = getenv("DTCSIM_DM_REG");
        o.reg_passes = !(e && atoi(e) == 0);
        e = getenv("DTCSIM_DM_TILE13");
        o.wide13 = !(e && atoi(e) == 0);
        return o;
    }();
    std::vector<DmStep> steps;
    {
        std::string err;
        if (!dm_plan(n, n_seg, seg_type, seg_off, q0, q1, val, probs, opt, steps, err)) return fail(DTC_ERR_INVALID, err.c_str());
    }
    double2* T = nullptr;
    CUDA_TRY(cudaMallocAsync((void**)&T, sizeof(double2) << n, s));
    int sweeps = 0;
    for (const DmStep& S : steps) {
        if (S.kind == 0) {
            k_dm_phase_table<<<((1 << n) + 127) / 128, 128, 0, s>>>(T, n, S.D);
        } else if (S.kind == 1) {
            const long long ne = 1ll << (2 * n);
            k_dm_apply_table<<<(unsigned)((ne + 255) / 256), 256, 0, s>>>((double2*)rho, n, T);
            ++sweeps;
        } else if (S.kind == 2) {
            const int TBg = S.P.tile_bits;
            if (TBg <= 12) k_dm_tile<256><<<1u << (2 * n - TBg), 256, sizeof(double2) << TBg, s>>>((double2*)rho, S.P, T);
            else k_dm_tile<512><<<1u << (2 * n - TBg), 512, sizeof(double2) << TBg, s>>>((double2*)rho, S.P, T);
            ++sweeps;
        } else {
            const int TBg = S.R.tile_bits;
            const size_t smem = S.R.n_rounds > 1 ? sizeof(double2) << TBg : 0;
            const unsigned grid = 1u << (2 * n - TBg);
            if (TBg == 12) k_dm_reg<256><<<grid, 256, smem, s>>>((double2*)rho, S.R, T);
            else k_dm_reg<512><<<grid, 512, smem, s>>>((double2*)rho, S.R, T);
            ++sweeps;
        }
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaFreeAsync(T, s));
    if (n_sweeps) *n_sweeps = sweeps;
    return DTC_OK;
}

int dtc_dm_probs(const void* rho, int n, int k, const int32_t* qubits, double* out, void* stream) {
    if (!rho || !out || n < 1 || n > 13 || k < 0 || k > n || (k > 0 && !qubits)) return fail(DTC_ERR_INVALID, "bad argument");
    DeviceGuard guard(device_of(rho));
    CUDA_TRY(guard.err);
    cudaStream_t s = (cudaStream_t)stream;
    int* dq = nullptr;
    CUDA_TRY(cudaMallocAsync(&dq, sizeof(int) * (k + 1), s));
    if (k) CUDA_TRY(cudaMemcpyAsync(dq, qubits, sizeof(int) * k, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(double) << k, s));
    const long long nr = 1ll << n;
    k_dm_probs<<<(unsigned)((nr + 127) / 128), 128, 0, s>>>((const double2*)rho, n, k, dq, out);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(s));
    cudaFreeAsync(dq, s);
    return DTC_OK;
}

static int shard_move(const void* state, void* buf, int n_local, int g, const int32_t* lq, int unpack, void* stream) {
    if (!state || !buf || g < 1 || g > 6 || g > n_local || !lq) return fail(DTC_ERR_INVALID, "bad argument");
    DeviceGuard guard(device_of(state));
    CUDA_TRY(guard.err);
    cudaStream_t s = (cudaStream_t)stream;
    int* dl = nullptr;
    CUDA_TRY(cudaMallocAsync(&dl, sizeof(int) * g, s));
    CUDA_TRY(cudaMemcpyAsync(dl, lq, sizeof(int) * g, cudaMemcpyHostToDevice, s));
    const long long ne = 1ll << n_local;
    k_shard_pack<<<(unsigned)((ne + 255) / 256), 256, 0, s>>>((const double2*)state, (double2*)buf, n_local, g, dl, unpack);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(s));
    cudaFreeAsync(dl, s);
    return DTC_OK;
}

int dtc_shard_pack(const void* state, void* out, int n_local, int g, const int32_t* lq, void* stream) {
    return shard_move(state, out, n_local, g, lq, 0, stream);
}
int dtc_shard_unpack(const void* in, void* state, int n_local, int g, const int32_t* lq, void* stream) {
    return shard_move(state, (void*)in, n_local, g, lq, 1, stream);
}

}  // extern "C"
