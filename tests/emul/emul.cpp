// emul.cpp -- CPU emulation of the dtcsim kernels (TEST INFRASTRUCTURE, never shipped or timed).
//
// Compiles the *same* host/device headers as the CUDA library (csrc/dtc_hd.cuh, csrc/dtc_core.hpp)
// with g++ and executes each CTA of k_tile_pass thread by thread, phase by phase (a phase ends where
// the kernel has a __syncthreads), so the tile-engine logic -- index maps, swizzles, table
// construction, sign masks, pass schedule -- is validated against the oracle without a GPU.
#include <stdlib.h>

#include <string>
#include <vector>

#include "../../noise-resilience-in-discrete-time-crystal-realizations-on-quantum-computers_b200/csrc/dtc_core.hpp"
#include "../../noise-resilience-in-discrete-time-crystal-realizations-on-quantum-computers_b200/csrc/dtc_readout.cuh"
#include "../../noise-resilience-in-discrete-time-crystal-realizations-on-quantum-computers_b200/csrc/dtc_dm.cuh"

template <int S2_LO>
static void emu_tile_pass(double2* state, const DtcTilePass& P, const DtcLayer* layers, const u64* masks,
                          long long n_traj, u64 rank_bits, u64 block, TileSmem& sm,
                          std::vector<double2>& regs) {
    const int ntb = P.n_local - DTC_TILE_BITS;
    const u64 tile = block & ((1ull << ntb) - 1);
    const u64 traj = block >> ntb;
    double2* st = state + (traj << P.n_local);
    const u64 base = tile_base_index(tile, P);
    const TileMasks M = tile_load_masks(P, masks, n_traj, traj);
    u64 off[DTC_THREADS];
    // --- up to the first barrier
    for (int tid = 0; tid < DTC_THREADS; ++tid) {
        double2* a = &regs[(size_t)tid * DTC_NREG];
        off[tid] = tile_thread_offset(tid, base, P);
        tile_gload(st, off[tid], P, a);
        if (P.layerD >= 0)
            tile_setup_thread(tid, sm, P, layers[P.layerD], base | (rank_bits << P.n_local), M.m1a, M.m1b, M.m2);
    }
    // --- second segment
    for (int tid = 0; tid < DTC_THREADS; ++tid) {
        double2* a = &regs[(size_t)tid * DTC_NREG];
        tile_rot_s1<S2_LO>(a, P.t1, P.tb, M.rmA);
        tile_sm_store13<S2_LO>(tid, sm, a);
        tile_tables_thread<S2_LO>(tid, sm, P);
    }
    // --- phase 2
    for (int tid = 0; tid < DTC_THREADS; ++tid) {
        double2 a[DTC_NREG];
        tile_sm_load2<S2_LO>(tid, sm, a);
        if (P.layerD >= 0 && P.nX > 0) tile_phase2_compute<S2_LO, true>(tid, a, sm, P, M.rmA, M.rmB);
        else tile_phase2_compute<S2_LO, false>(tid, a, sm, P, M.rmA, M.rmB);
        // a thread only rewrites the slots it read, so doing this without a barrier is faithful
        tile_sm_store2<S2_LO>(tid, sm, a);
    }
    // --- phase 3
    for (int tid = 0; tid < DTC_THREADS; ++tid) {
        double2 a[DTC_NREG];
        tile_sm_load13<S2_LO>(tid, sm, a);
        tile_rot_s1<S2_LO>(a, P.t2, P.tb, M.rmB);
        tile_gstore(st, off[tid], P, a);
    }
}


// k_tile_stream: the TMA load / store are emulated by a gather / scatter of the tile into a dense stage
// buffer; the __syncwarp()s of the table-builder warp and the phase boundaries are loop boundaries.
template <int MODE>
static void emu_stream_pass(double2* state, const DtcStreamPass& P, const DtcLayer* layers, const u64* masks,
                            long long n_traj, u64 rank_bits, u64 T, std::vector<double2>& stage, StreamSlot& tab, StreamBuild& bl) {
    const int ntb = P.n_local - DTC_TILE_BITS;
    const u64 traj = T >> ntb, tit = T & ((1ull << ntb) - 1);
    double2* st = state + (traj << P.n_local);
    const u64 base = stream_tile_base(tit, P);
    auto gidx = [&](int l) {
        u64 o = base;
        for (int b = 0; b < DTC_TILE_BITS; ++b)
            if ((l >> b) & 1) o += 1ull << P.tb[b];
        return o;
    };
    for (int l = 0; l < DTC_TILE; ++l) stage[l] = st[gidx(l)];
    const StreamMasks M = stream_load_masks(P, masks, n_traj, traj);
    // table-builder warp: three steps separated by __syncwarp()
    if (P.layerD >= 0) {
        const DtcLayer& L = layers[P.layerD];
        for (int lane = 0; lane < 32; ++lane) stream_build1(lane, bl, P, L, base | (rank_bits << P.n_local), M.m1a, M.m1b, M.m2);
        for (int lane = 0; lane < 32; ++lane) stream_build2(lane, bl, P, L);
    }
    for (int lane = 0; lane < 32; ++lane) stream_build3(lane, bl, tab, P);
    tab.rmA = M.rmA; tab.rmB = M.rmB;
    if (MODE == 3) {
        for (int t = 0; t < 128; ++t) stream_phaseC(t, stage.data(), tab, P, M.rmA, M.rmB);
    } else {
        constexpr int M13 = MODE == 3 ? 1 : MODE;
        if (P.layerA >= 0)
            for (int t = 0; t < 128; ++t) stream_phase13<M13>(t, stage.data(), P.t1, P.tb, M.rmA);
        const bool all3 = P.layerA >= 0 && P.layerD >= 0 && P.layerB >= 0;
        for (int t = 0; t < 128; ++t) {
            if (all3) stream_phase2(t, stage.data(), tab, P, M.rmA, M.rmB);
            else stream_phase2_partial(t, stage.data(), tab, P, M.rmA, M.rmB);
        }
        if (P.layerB >= 0)
            for (int t = 0; t < 128; ++t) stream_phase13<M13>(t, stage.data(), P.t2, P.tb, M.rmB);
    }
    for (int l = 0; l < DTC_TILE; ++l) st[gidx(l)] = stage[l];
}

extern "C" int emu_run(int n_qubits, int n_layers, int64_t n_events, const int32_t* type, const int32_t* layer,
                       const int32_t* q0, const int32_t* q1, const int32_t* slot, const double* val,
                       const double* probs, double global_phase, int engine, int n_local, int n_exec_layers, int64_t n_traj,
                       int64_t traj_offset, u64 seed, u64 init_index, u64 rank_bits, double* state_out,
                       u64* fx, u64* fz, int* ph, int* n_passes, char* errbuf, int errlen) {
    DtcProgramHost P;
    P.n_qubits = n_qubits;
    P.n_layers = n_layers;
    P.n_exec_layers = n_exec_layers > 0 ? n_exec_layers : n_layers;
    P.n_local = n_local;
    std::string err;
    auto bail = [&](const std::string& m) {
        snprintf(errbuf, errlen, "%s", m.c_str());
        return -1;
    };
    if (!dtc_stage_events(P, n_events, type, layer, q0, q1, slot, val, probs, global_phase, err)) return bail(err);
    if (!dtc_build_layers(P, err)) return bail(err);
    if (engine == 0) engine = (n_local >= DTC_TILE_BITS) ? 2 : 1;
    const bool use_stream = engine == 3;      // 3: streaming engine where a pass is eligible, tile engine elsewhere
    if (engine == 3) engine = 2;
    int n_stream = 0;
    if (engine == 2) {
        if (n_local < DTC_TILE_BITS) return bail("tile engine needs n_local >= 12");
        if (!dtc_schedule_tile(P, err)) return bail(err);
        *n_passes = (int)P.passes.size();
    } else {
        dtc_schedule_generic(P);
        *n_passes = (int)P.gsteps.size();
    }
    std::vector<u64> masks((size_t)n_layers * 4 * n_traj, 0);
    for (int64_t t = 0; t < n_traj; ++t)
        frame_walk((u64)(traj_offset + t), seed, P.events.data(), (int64_t)P.events.size(), masks.data() + t, n_traj,
                   fx + t, fz + t, ph + t);
    double2* st = reinterpret_cast<double2*>(state_out);
    const size_t ne = (size_t)n_traj << n_local;
    for (size_t i = 0; i < ne; ++i) st[i] = make_double2(0.0, 0.0);
    for (int64_t t = 0; t < n_traj; ++t) st[((u64)t << n_local) + init_index] = make_double2(1.0, 0.0);
    if (engine == 2) {
        TileSmem* sm = new TileSmem();
        std::vector<double2> regs((size_t)DTC_THREADS * DTC_NREG);
        const u64 grid = (u64)n_traj << (n_local - DTC_TILE_BITS);
        std::vector<double2> stage(DTC_TILE);
        StreamSlot* tab = new StreamSlot();
        StreamBuild* bl = new StreamBuild();
        for (size_t ip = 0; ip < P.passes.size(); ++ip) {
            const DtcTilePass& T = P.passes[ip];
            const DtcStreamPass& S = P.spasses[ip];
            if (use_stream && S.mode) {
                ++n_stream;
                for (u64 b = 0; b < grid; ++b) {
                    if (S.mode == 1) emu_stream_pass<1>(st, S, P.layers.data(), masks.data(), n_traj, rank_bits, b, stage, *tab, *bl);
                    else if (S.mode == 2) emu_stream_pass<2>(st, S, P.layers.data(), masks.data(), n_traj, rank_bits, b, stage, *tab, *bl);
                    else emu_stream_pass<3>(st, S, P.layers.data(), masks.data(), n_traj, rank_bits, b, stage, *tab, *bl);
                }
                continue;
            }
            for (u64 b = 0; b < grid; ++b) {
                if (T.s2_lo == 0) emu_tile_pass<0>(st, T, P.layers.data(), masks.data(), n_traj, rank_bits, b, *sm, regs);
                else if (T.s2_lo == 1) emu_tile_pass<1>(st, T, P.layers.data(), masks.data(), n_traj, rank_bits, b, *sm, regs);
                else emu_tile_pass<2>(st, T, P.layers.data(), masks.data(), n_traj, rank_bits, b, *sm, regs);
            }
        }
        delete sm;
        delete tab;
        delete bl;
        if (use_stream) *n_passes = n_stream;      // callers of engine 3 want to know how many passes streamed
    } else {
        for (const DtcGenericStep& g : P.gsteps) {
            if (g.kind == 0) {
                const u64 np = (u64)n_traj << (n_local - 1);
                for (u64 i = 0; i < np; ++i) {
                    const u64 traj = i >> (n_local - 1);
                    const u64 p = i & ((1ull << (n_local - 1)) - 1);
                    const u64 lowm = (1ull << g.q) - 1;
                    const u64 i0 = ((p & ~lowm) << 1) | (p & lowm);
                    double2* s2 = st + (traj << n_local);
                    const u64 rm = masks[(size_t)(g.layer * 4) * n_traj + traj];
                    const double ts = ((rm >> g.q) & 1ull) ? -g.t : g.t;
                    rot_pair(s2[i0], s2[i0 | (1ull << g.q)], ts);
                }
            } else {
                const DtcLayer& L = P.layers[g.layer];
                for (u64 i = 0; i < ne; ++i) {
                    const u64 traj = i >> n_local, x = i & ((1ull << n_local) - 1);
                    const u64* m = masks.data() + (size_t)(g.layer * 4) * n_traj;
                    st[i] = cmul(st[i], diag_phase(L, x | (rank_bits << n_local), m[1 * n_traj + traj],
                                                   m[2 * n_traj + traj], m[3 * n_traj + traj]));
                }
            }
        }
    }
    return 0;
}

// expose the pass schedule for inspection by tests: fills up to `cap` rows of
// [s2_lo, layerA, layerD, layerB, nT1, nT2, nX, nC, nO, tb0..tb11, stream mode (0: not eligible)]
extern "C" int emu_schedule(int n_qubits, int n_layers, int64_t n_events, const int32_t* type, const int32_t* layer,
                            const int32_t* q0, const int32_t* q1, const int32_t* slot, const double* val,
                            const double* probs, int n_local, int n_exec_layers, int32_t* rows, int cap, char* errbuf, int errlen) {
    DtcProgramHost P;
    P.n_qubits = n_qubits;
    P.n_layers = n_layers;
    P.n_exec_layers = n_exec_layers > 0 ? n_exec_layers : n_layers;
    P.n_local = n_local;
    std::string err;
    if (!dtc_stage_events(P, n_events, type, layer, q0, q1, slot, val, probs, 0.0, err) || !dtc_build_layers(P, err) ||
        !dtc_schedule_tile(P, err)) {
        snprintf(errbuf, errlen, "%s", err.c_str());
        return -1;
    }
    int n = 0;
    for (const DtcTilePass& T : P.passes) {
        if (n >= cap) break;
        int32_t* r = rows + (size_t)n * 22;
        r[0] = T.s2_lo; r[1] = T.layerA; r[2] = T.layerD; r[3] = T.layerB;
        r[4] = T.nT1; r[5] = T.nT2; r[6] = T.nX; r[7] = T.nC; r[8] = T.nO;
        for (int l = 0; l < 12; ++l) r[9 + l] = T.tb[l];
        r[21] = P.spasses[(size_t)n].mode;
        ++n;
    }
    return (int)P.passes.size();
}

// k_readout_small: the per-trajectory read-out function of csrc/dtc_readout.cuh on the CPU.
// masks: [n_layers*4][n_traj] as written by frame_walk; rdm: [n_traj][2^n_reg][2^n_reg]; probs: [n_traj][2^m].
extern "C" int emu_readout_small(int n_qubits, int n_layers, int64_t n_events, const int32_t* type, const int32_t* layer,
                                 const int32_t* q0, const int32_t* q1, const int32_t* slot, const double* val,
                                 const double* probs_in, int64_t n_small, const int64_t* small_events, int n_reg,
                                 const int32_t* reg_bits, int n_elim, const int32_t* elim_bits, int m,
                                 const int32_t* measure_bits, const double* rdm, const u64* masks, const u64* fx,
                                 int64_t n_traj, double* probs_out, char* errbuf, int errlen) {
    DtcProgramHost P;
    P.n_qubits = n_qubits;
    P.n_layers = n_layers;
    P.n_exec_layers = n_layers;
    P.n_local = n_qubits;
    std::string err;
    if (!dtc_stage_events(P, n_events, type, layer, q0, q1, slot, val, probs_in, 0.0, err) || !dtc_build_layers(P, err)) {
        snprintf(errbuf, errlen, "%s", err.c_str());
        return -1;
    }
    DtcSmallPlan S;
    memset(&S, 0, sizeof(S));
    S.nq = n_reg + n_elim; S.n_reg = n_reg; S.m = m;
    for (int i = 0; i < n_reg; ++i) S.bits[i] = reg_bits[i];
    for (int i = 0; i < n_elim; ++i) S.bits[n_reg + i] = elim_bits[i];
    for (int i = 0; i < m; ++i) { S.meas_bit[i] = measure_bits[i]; S.meas_pos[i] = small_pos(S, measure_bits[i]); }
    std::vector<long long> idx(small_events, small_events + n_small);
    const double2* r = reinterpret_cast<const double2*>(rdm);
    for (int64_t t = 0; t < n_traj; ++t)
        small_readout_traj(S, P.events.data(), idx.data(), n_small, r + (t << (2 * n_reg)), masks + t, n_traj, fx[t],
                           probs_out + (t << m));
    return 0;
}

extern "C" void emu_set_high_stride_bit(int bit) { g_dtc_high_stride_bit = bit; }

// dtc_dm_run on the CPU: the planner of csrc/dtc_dm.cuh, then every step executed CTA by CTA.  Register passes (k_dm_reg) run
// the kernel's own per-thread function round by round (a round ends where the kernel has its __syncthreads); element-per-thread
// tile passes (k_dm_tile) are replayed from their pass descriptor.  info[0] = sweeps over rho, info[1] = register passes,
// info[2] = worst shared-memory conflict degree of a quarter warp over all register-pass rounds (1 = conflict free).
extern "C" int emu_dm_run(int n, int n_seg, const int32_t* seg_type, const int32_t* seg_off, const int32_t* q0, const int32_t* q1,
                          const double* val, const double* probs, int reg_passes, int wide13, double* rho_io, int32_t* info,
                          char* errbuf, int errlen) {
    DmPlanOptions opt;
    opt.reg_passes = reg_passes != 0;
    opt.wide13 = wide13 != 0;
    std::vector<DmStep> steps;
    std::string err;
    if (!dm_plan(n, n_seg, seg_type, seg_off, q0, q1, val, probs, opt, steps, err)) {
        snprintf(errbuf, errlen, "%s", err.c_str());
        return -1;
    }
    double2* rho = reinterpret_cast<double2*>(rho_io);
    const u64 ne = 1ull << (2 * n), rmask = (1ull << n) - 1;
    std::vector<double2> T((size_t)1 << n);
    int sweeps = 0, nreg = 0, worst = 1;
    for (const DmStep& S : steps) {
        if (S.kind == 0) {
            for (u64 x = 0; x < (1ull << n); ++x) {
                double ang = 0.0;
                for (int k = 0; k < S.D.n1; ++k) ang += 0.5 * S.D.a[k] * (double)(1 - 2 * (int)((x >> S.D.q1[k]) & 1));
                for (int k = 0; k < S.D.n2; ++k) ang += 0.5 * S.D.b[k] * (double)(1 - 2 * (int)(((x >> S.D.qi[k]) ^ (x >> S.D.qj[k])) & 1));
                T[x] = make_double2(cos(ang), -sin(ang));
            }
        } else if (S.kind == 1) {
            for (u64 i = 0; i < ne; ++i) {
                const double2 tc = T[i >> n];
                rho[i] = cmul(rho[i], cmul(T[i & rmask], make_double2(tc.x, -tc.y)));
            }
            ++sweeps;
        } else if (S.kind == 2) {
            const DmTilePass& P = S.P;
            const int te = 1 << P.tile_bits;
            std::vector<double2> tile(te);
            std::vector<u64> off(te);
            for (int e = 0; e < te; ++e) {
                u64 o = 0;
                for (int l = 0; l < P.tile_bits; ++l)
                    if ((e >> l) & 1) o |= 1ull << P.tb[l];
                off[e] = o;
            }
            for (u64 cta = 0; cta < (ne >> P.tile_bits); ++cta) {
                const u64 base = dm_cta_base(cta, P.seg_n, P.seg_src, P.seg_len, P.seg_dst);
                for (int e = 0; e < te; ++e) {
                    const u64 g = base | off[e];
                    double2 v = rho[g];
                    if (P.has_diag) {
                        const double2 tc = T[g >> n];
                        v = cmul(v, cmul(T[g & rmask], make_double2(tc.x, -tc.y)));
                    }
                    tile[e] = v;
                }
                for (int k = 0; k < P.nq; ++k) {
                    const DmQubitOp& Q = P.q[k];
                    const double c = Q.c, s = Q.s;
                    for (int x = 0; x < te; ++x) {
                        if (x & ((1 << Q.lr) | (1 << Q.lc))) continue;
                        const int i00 = x, i10 = x | (1 << Q.lr), i01 = x | (1 << Q.lc), i11 = i10 | i01;
                        double2 e00 = tile[i00], e10 = tile[i10], e01 = tile[i01], e11 = tile[i11];
                        const double2 a00 = make_double2(c * e00.x + s * e10.y, c * e00.y - s * e10.x);
                        const double2 a10 = make_double2(c * e10.x + s * e00.y, c * e10.y - s * e00.x);
                        const double2 a01 = make_double2(c * e01.x + s * e11.y, c * e01.y - s * e11.x);
                        const double2 a11 = make_double2(c * e11.x + s * e01.y, c * e11.y - s * e01.x);
                        e00 = make_double2(c * a00.x - s * a01.y, c * a00.y + s * a01.x);
                        e01 = make_double2(c * a01.x - s * a00.y, c * a01.y + s * a00.x);
                        e10 = make_double2(c * a10.x - s * a11.y, c * a10.y + s * a11.x);
                        e11 = make_double2(c * a11.x - s * a10.y, c * a11.y + s * a10.x);
                        tile[i00] = make_double2(Q.dA * e00.x + Q.dB * e11.x, Q.dA * e00.y + Q.dB * e11.y);
                        tile[i11] = make_double2(Q.dA * e11.x + Q.dB * e00.x, Q.dA * e11.y + Q.dB * e00.y);
                        tile[i10] = make_double2(Q.oA * e10.x + Q.oB * e01.x, Q.oA * e10.y + Q.oB * e01.y);
                        tile[i01] = make_double2(Q.oA * e01.x + Q.oB * e10.x, Q.oA * e01.y + Q.oB * e10.y);
                    }
                }
                for (int e = 0; e < te; ++e) rho[base | off[e]] = tile[e];
            }
            ++sweeps;
        } else {
            const DmRegPass& P = S.R;
            const int threads = 1 << (P.tile_bits - 4);
            std::vector<double2> tile((size_t)1 << P.tile_bits);
            // (CTA index, tile element) -> global index must be a bijection onto [0, 4^n): the tile bits and the bits the
            // CTA-index segments land on are disjoint and together are exactly the 2n index bits (no out-of-range access)
            {
                u64 cover = 0;
                bool ok = true;
                int src_bits = 0;
                for (int l = 0; l < P.tile_bits; ++l) {
                    const u64 b = 1ull << P.tb[l];
                    ok = ok && P.tb[l] >= 0 && P.tb[l] < 2 * n && !(cover & b);
                    cover |= b;
                }
                for (int k = 0; k < P.seg_n; ++k) {
                    ok = ok && P.seg_src[k] == src_bits;
                    for (int j = 0; j < P.seg_len[k]; ++j) {
                        const u64 b = 1ull << (P.seg_dst[k] + j);
                        ok = ok && P.seg_dst[k] + j < 2 * n && !(cover & b);
                        cover |= b;
                    }
                    src_bits += P.seg_len[k];
                }
                if (!ok || cover != ne - 1 || src_bits != 2 * n - P.tile_bits) {
                    snprintf(errbuf, errlen, "register pass: tile bits and CTA-index segments do not partition the index bits");
                    return -1;
                }
                for (int r = 0; r < P.n_rounds; ++r)
                    for (int k = 0; k < 4; ++k)
                        if (P.r[r].hg[k] != (1u << P.tb[P.r[r].hb[k]]) || P.r[r].hp[k] != dm_phys(1 << P.r[r].hb[k])) {
                            snprintf(errbuf, errlen, "register pass: derived masks of round %d are inconsistent", r);
                            return -1;
                        }
            }
            // every element of the tile must belong to exactly one (thread, register) of every round
            for (int r = 0; r < P.n_rounds; ++r) {
                std::vector<int> seen((size_t)1 << P.tile_bits, 0);
                for (int tid = 0; tid < threads; ++tid) {
                    int e = 0;
                    for (int k = 0; k < P.tile_bits - 4; ++k)
                        if ((tid >> k) & 1) e |= 1 << P.r[r].tbit[k];
                    for (int j = 0; j < 16; ++j) {
                        int ej = e;
                        for (int k = 0; k < 4; ++k)
                            if ((j >> k) & 1) ej |= 1 << P.r[r].hb[k];
                        ++seen[ej];
                    }
                }
                for (int x : seen)
                    if (x != 1) {
                        snprintf(errbuf, errlen, "register pass: round %d does not cover the tile exactly once", r);
                        return -1;
                    }
                // conflict degree of the 16 B accesses of each quarter warp (same register j, eight consecutive lanes)
                for (int q8 = 0; q8 < threads / 8; ++q8) {
                    int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    for (int lane = 0; lane < 8; ++lane) {
                        const int tid = q8 * 8 + lane;
                        int e = 0;
                        for (int k = 0; k < P.tile_bits - 4; ++k)
                            if ((tid >> k) & 1) e |= 1 << P.r[r].tbit[k];
                        ++cnt[dm_phys(e) & 7];
                    }
                    for (int b = 0; b < 8; ++b)
                        if (cnt[b] > worst) worst = cnt[b];
                }
            }
            for (u64 cta = 0; cta < (ne >> P.tile_bits); ++cta) {
                const u64 base = dm_cta_base(cta, P.seg_n, P.seg_src, P.seg_len, P.seg_dst);
                for (int r = 0; r < P.n_rounds; ++r)
                    for (int tid = 0; tid < threads; ++tid) {
                        if (P.tile_bits == 12) dmr_round<8>(r, tid, P, rho, T.data(), (unsigned)base, tile.data());
                        else dmr_round<9>(r, tid, P, rho, T.data(), (unsigned)base, tile.data());
                    }
            }
            ++sweeps;
            ++nreg;
        }
    }
    info[0] = sweeps; info[1] = nreg; info[2] = worst;
    return 0;
}

// k_dm_superop on the CPU: superop = 16 complex numbers, row major, (re, im) pairs
extern "C" int emu_dm_superop(int n, int q, const double* superop, double* rho_io) {
    DmSuperop S;
    for (int k = 0; k < 16; ++k) { S.re[k] = superop[2 * k]; S.im[k] = superop[2 * k + 1]; }
    double2* rho = reinterpret_cast<double2*>(rho_io);
    for (long long i = 0; i < (1ll << (2 * n - 2)); ++i) dm_superop_thread(rho, n, q, S, i);
    return 0;
}
