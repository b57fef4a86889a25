/* dtcsim C ABI -- the drop-in boundary of the B200-native DTC Floquet-circuit simulator.
 *
 * What it replaces.  The reference reaches its simulator through Python, not FFI:
 *     backend = AerSimulator(noise_model=nm, device="GPU", cuStateVec_enable=True)   fast.py:156
 *     result  = backend.run(circ_tnoise, shots=1024).result()                          fast.py:211
 *     counts  = result.get_counts(circ_tnoise)                                         fast.py:212
 * Everything below `run()` is qiskit-aer C++/CUDA (third party, not in the reference tree).  This
 * header is the C boundary our Python host (package module backend.py, class DTCSimulator) binds
 * with ctypes; a maintainer of the reference would bind the same symbols.
 *
 * Conventions.  Plain pointers and sizes only.  Device buffers are caller-owned (allocated by the
 * host runtime, e.g. torch) and never freed here; host arrays are consumed before the call
 * returns.  Every function returns 0 on success or a negative dtc_status and records a message
 * retrievable with dtc_last_error() (thread-local).  `stream` is a cudaStream_t passed as void*.
 * States are complex128, interleaved (re, im), basis index little-endian (qubit k = bit k), one
 * state of 2^n amplitudes per trajectory, trajectories contiguous.
 *
 * Execution model.  A circuit is compiled on the host into an event list of
 *   ROT(q, theta) = exp(-i theta X_q/2),  D1(q, a) = exp(-i a Z_q/2),  D2(i,j,b) = exp(-i b Z_i Z_j/2),
 *   NOISE(q, pX,pY,pZ)
 * grouped in layers  D_0 | R_1 D_1 | R_2 D_2 ...  A sampled Pauli never touches the state: it
 * updates a per-trajectory Pauli frame F (psi_true = F psi'), which only flips signs of later
 * angles.  dtc_program_run() samples the frames (Philox4x32-10), then streams the batch of states
 * once per fused  R|S -> D -> R|S  pass.
 */
#ifndef DTCSIM_H
#define DTCSIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    DTC_OK = 0,
    DTC_ERR_INVALID = -1,   /* bad argument / inconsistent program */
    DTC_ERR_CUDA = -2,      /* CUDA runtime error (message has the CUDA string) */
    DTC_ERR_UNSUPPORTED = -3,
    DTC_ERR_NOMEM = -4
} dtc_status;

enum { DTC_EV_ROT = 0, DTC_EV_D1 = 1, DTC_EV_D2 = 2, DTC_EV_NOISE = 3, DTC_EV_D2C = 4 };
enum { DTC_ENGINE_AUTO = 0, DTC_ENGINE_GENERIC = 1, DTC_ENGINE_TILE = 2 };
/* special init_index values of dtc_program_run: continue from the buffer's contents / start from the zero vector
 * (shards of a distributed state that do not hold the basis amplitude) */
#define DTC_INIT_KEEP 0xFFFFFFFFFFFFFFFFull
#define DTC_INIT_ZERO 0xFFFFFFFFFFFFFFFEull

typedef struct dtc_program dtc_program;   /* opaque compiled circuit */

/* ---- library ------------------------------------------------------------------------------ */
int dtc_version(void);                     /* 100*major + minor */
const char *dtc_last_error(void);          /* replaces Python exceptions raised inside Aer's run() */
int dtc_device_count(int *count);

/* ---- program: the compiled form of one circuit (replaces Aer's internal circuit/op list that
 *      backend.run() builds from the QuantumCircuit, fast.py:211) ------------------------------ */
/* n_qubits: qubits of the full register (global index width); n_local: qubits held by this
 * process' state (n_local == n_qubits unless the state is sharded on its top bits). */
int dtc_program_create(int n_qubits, int n_layers, dtc_program **out);
int dtc_program_destroy(dtc_program *p);
/* Event arrays in circuit order (see plan.py): type[n], layer[n], q0[n], q1[n], slot[n]
 * (D1: sign slot 0/1; D2: term index in its layer; NOISE: site id), val[n] (ROT theta, D1 a,
 * D2 b), probs[n][3] (NOISE pX,pY,pZ).  global_phase: psi_true carries exp(+i global_phase). */
int dtc_program_set_events(dtc_program *p, int64_t n_events, const int32_t *type, const int32_t *layer,
                           const int32_t *q0, const int32_t *q1, const int32_t *slot,
                           const double *val, const double *probs, double global_phase);
/* Only layers [0, n_exec_layers) are executed on the state; later ("virtual") layers still get sign masks
 * from the frame walk -- the host uses them for the small read-out simulation (plan.py, optimize=True).
 * DTC_EV_D2C is a D2 term whose partner q1 is still |0> in psi': sign from (q0,q1), phase one-body on q0. */
int dtc_program_set_exec_layers(dtc_program *p, int n_exec_layers);
/* Optional, before finalize: the internal bit index of the single read-out qubit of a factorised circuit.  The pass
 * schedule is then chosen (among equally long ones) so that its last pass works on a tile holding that qubit, which
 * lets that pass fuse the read-out reduction (dtc_program_set_fused_rdm). */
int dtc_program_set_readout_hint(dtc_program *p, int bit);
/* Build layer tables and the pass schedule and upload them to `device`.
 * engine: DTC_ENGINE_AUTO picks the fused tile engine when n_local >= 12. */
int dtc_program_finalize(dtc_program *p, int device, int engine, int n_local);
int dtc_program_num_passes(const dtc_program *p, int *n_passes);
/* bytes of caller-owned device scratch dtc_program_run() needs for n_traj trajectories */
int dtc_program_workspace_bytes(const dtc_program *p, int64_t n_traj, size_t *bytes);
/* Evolve n_traj trajectories (global ids traj_offset .. traj_offset+n_traj-1) from the basis state
 * `init_index` (per process: local index; rank_bits are the value of the global index bits above
 * n_local).  state: device buffer of n_traj * 2^n_local complex128.  On return the buffer holds
 * psi' (frame not applied); the final frames live in the workspace (dtc_program_frames).
 * Replaces the per-shot statevector evolution inside AerSimulator.run() (fast.py:211). */
int dtc_program_run(dtc_program *p, void *state, int64_t n_traj, int64_t traj_offset, uint64_t seed,
                    uint64_t init_index, uint64_t rank_bits, void *workspace, size_t workspace_bytes,
                    void *stream);
/* dtc_program_run in two steps, for callers that interleave the passes of a program with other work (sharded.py: a
 * shard is swept slice by slice and each finished slice is sent while the next one is swept).
 * prepare(): the frame walk only (sign masks + final frames into the workspace).  run_passes(): passes [pass_begin, pass_end)
 * of the tile engine's schedule on `state`, using the masks prepare() left in the workspace; init_index other than
 * DTC_INIT_KEEP is allowed when pass_begin == 0.  store_last_or_null: the LAST pass of the range stores its tiles to that
 * buffer (same layout as `state`) instead of in place -- it may be a peer GPU's memory mapped into this process, so the
 * sweep's TMA stores travel over NVLink straight into the receiver's buffer (the pass must be a streaming pass with
 * contiguous tiles: dtc_program_pass_info).  n_ctas > 0 limits the persistent grid, leaving SMs to a kernel on another
 * stream. */
int dtc_program_prepare(dtc_program *p, int64_t n_traj, int64_t traj_offset, uint64_t seed, void *workspace,
                        size_t workspace_bytes, void *stream);
int dtc_program_run_passes(dtc_program *p, void *state, void *store_last_or_null, int pass_begin, int pass_end, int n_ctas,
                           int64_t n_traj, uint64_t init_index, uint64_t rank_bits, void *workspace,
                           size_t workspace_bytes, void *stream);
int dtc_program_pass_info(const dtc_program *p, int pass, int *streaming, int *contiguous);
/* Device pointers (inside the workspace) to the final frame of each trajectory:
 * fx, fz: uint64[n_traj] bit masks; ph: int32[n_traj] power of i.  psi_true = i^ph X^fx Z^fz psi'. */
int dtc_program_frames(const dtc_program *p, void *workspace, int64_t n_traj,
                       uint64_t **fx, uint64_t **fz, int32_t **ph);

/* Engine selection inside the tile engine.  A fused pass whose tile is [0,12) or {0,1}+[g,g+10) runs on the
 * TMA-fed persistent kernel k_tile_stream; other passes (and all passes after dtc_set_stream_engine(0)) run on
 * the register-fed k_tile_pass.  enable < 0 restores the default (environment DTCSIM_STREAM, else on).
 * dtc_program_num_stream_passes(): how many passes of the schedule are eligible for k_tile_stream. */
int dtc_set_stream_engine(int enable);
/* Qubit groups that start at or above this internal bit (default 15: 64 B runs of such a group's tiles would each lie in a
 * different 2 MiB page) are given five qubits and tiles of 32 runs of 2 KB.  Affects programs finalized afterwards;
 * tests lower it to exercise that path on small registers. */
int dtc_set_high_stride_bit(int bit);
int dtc_program_num_stream_passes(const dtc_program *p, int *n_passes);
/* Persistent CTAs k_tile_stream launches (default 0 = one per SM).  A smaller grid leaves SMs to kernels that must run
 * at the same time on other streams (NCCL send/recv while a sharded state is exchanged, sharded.py). */
int dtc_set_stream_ctas(int n_ctas);

/* Kernel timing for roofline accounting: when enabled, dtc_program_run() brackets its pass loop with
 * CUDA events on the launching stream; dtc_program_pass_time() waits for the last run and returns the
 * elapsed milliseconds and the number of state-sweep launches in that run. */
int dtc_program_set_profiling(dtc_program *p, int enable);
/* Traffic accounting of the last dtc_program_run: *gen_first = 1 if its first pass generated the initial state
 * (wrote the batch, read nothing), *fused_rdm = 1 if its last pass reduced the read-out density matrix instead
 * of storing (read the batch, wrote nothing). */
int dtc_program_last_run_flags(const dtc_program *p, int *gen_first, int *fused_rdm);
int dtc_program_pass_time(dtc_program *p, float *ms, int *n_launches);
/* Per-pass durations of the last whole-program run with profiling on: ms[i], and modes[i] = tile layout of pass i
 * (1 contiguous 64 KB tiles, 2 runs of 64 B, 3 runs of 2 KB, 0 register-fed kernel); *n_passes = passes in that run. */
int dtc_program_pass_times(dtc_program *p, float *ms, int *modes, int cap, int *n_passes);

/* Read-out of a factorised circuit (the Hadamard-test ancilla kept out of the register; replaces the measure
 * sampling input of AerSimulator.run(), fast.py:211).  set_readout() (once, after finalize): the indices of the
 * "small" events in the event list, the internal bit indices of the register qubits whose reduced density
 * matrix is read back (<= 2), of the eliminated qubits, and of the measured qubits (1..3, all inside those).
 * readout(): rdm = device output of dtc_rdm on reg_bits for the same n_traj; workspace = the one dtc_program_run
 * used (sign masks, frames); probs[n_traj][2^m] (device, double), bit i of the column = i-th measured bit. */
int dtc_program_set_readout(dtc_program *p, int64_t n_small, const int64_t *small_events, int n_reg,
                            const int32_t *reg_bits, int n_elim, const int32_t *elim_bits, int m,
                            const int32_t *measure_bits);
int dtc_program_readout(const dtc_program *p, const void *rdm, void *workspace, int64_t n_traj, double *probs,
                        void *stream);
/* Fused read-out: when enabled and the program's last pass runs on k_tile_stream with the (single) read-out qubit
 * inside its tile, that pass accumulates the qubit's reduced density matrix into the workspace instead of storing
 * the state (one read + one write of the batch saved per circuit; the state buffer then holds psi' BEFORE the last
 * pass and must not be used).  *active_or_null tells whether the next dtc_program_run will fuse;
 * dtc_program_fused_rdm() returns the device pointer [n_traj][2][2] complex128 of the last run (NULL if not fused). */
int dtc_program_set_fused_rdm(dtc_program *p, int enable, int *active_or_null);
int dtc_program_fused_rdm(const dtc_program *p, void *workspace, int64_t n_traj, void **rdm);

/* Resident execution (read-out-only runs; replaces the whole per-shot evolution loop of AerSimulator.run(), fast.py:211,
 * in ONE persistent launch).  When every pass of the program runs on the streaming engine and the fused read-out is active
 * (dtc_program_set_fused_rdm), the trajectories are processed in groups of `group` that go through ALL passes before the
 * next group starts and share `group` state slots, so the sweeps read and write L2-resident data instead of streaming the
 * batch through HBM once per pass.  resident_info(): eligibility, group size and the scratch bytes (group x 2^n_local
 * complex128) run_resident() needs instead of a state buffer for the whole batch.  The result is the fused read-out density
 * matrix (dtc_program_fused_rdm) and the frames, exactly as after dtc_program_run with the fused read-out; the scratch
 * buffer holds no defined state afterwards.  Returns DTC_ERR_UNSUPPORTED when the program is not eligible.
 * dtc_set_resident_bytes(): state bytes kept in flight per group (default 64 MiB of the 126 MB L2). */
int dtc_set_resident_bytes(size_t bytes);
int dtc_program_resident_info(const dtc_program *p, int64_t n_traj, int *eligible, int *group, size_t *scratch_bytes);
int dtc_program_run_resident(dtc_program *p, void *scratch, size_t scratch_bytes, int64_t n_traj, int64_t traj_offset,
                             uint64_t seed, uint64_t init_index, uint64_t rank_bits, void *workspace,
                             size_t workspace_bytes, void *stream);
/* *resident = 1 if the last run used resident execution; *kernel_launches = kernels that run launched. */
int dtc_program_last_run_info(const dtc_program *p, int *resident, int *kernel_launches);

/* ---- state utilities ---------------------------------------------------------------------- */
/* In-place psi' -> psi_true for each trajectory (used for amplitude-level parity / save_statevector). */
int dtc_materialize(void *state, int n_local, int64_t n_traj, const uint64_t *fx, const uint64_t *fz,
                    const int32_t *ph, void *scratch_one_state, void *stream);
/* Marginal probabilities of `k` (<= 12) qubits: out[n_traj][2^k] (device, double), bit i of the
 * column index = qubit qubits[i]; frame x-bits (may be NULL) flip the outcome of their qubit.
 * Replaces the measure sampling input of Aer (fast.py:211) and compute_z_expectation's p0/p1. */
int dtc_probs(const void *state, int n_local, int64_t n_traj, int k, const int32_t *qubits,
              const uint64_t *fx_or_null, double *out, void *stream);
/* Reduced density matrix of k (<= 2) qubits of psi': out[n_traj][2^k][2^k] complex128 (device, row-major,
 * rho[a][b] = sum_rest psi(a,rest) conj psi(b,rest); bit i of a = qubit qubits[i]).  Read-out of the
 * Hadamard-test signal without carrying the ancilla in the register (SURVEY.md 8a identity). */
int dtc_rdm(const void *state, int n_local, int64_t n_traj, int k, const int32_t *qubits, void *out, void *stream);
/* <Z_q> for every qubit: out[n_traj][n_local] (device, double).  (dtc_qasm.py:145 per-qubit <Z_i>) */
int dtc_expect_z(const void *state, int n_local, int64_t n_traj, const uint64_t *fx_or_null,
                 double *out, void *stream);
/* Inverse-CDF sampling from rows of a probability table (device): row r of probs[n_rows][n_cols];
 * sample s of row r uses u = philox(seed; index = s, stream 1, traj = traj_offset + r) and writes
 * out[r * n_samples + s] (int32 column index).  Replaces Aer's measurement sampler. */
int dtc_sample_rows(const double *probs, int64_t n_rows, int n_cols, int n_samples, uint64_t seed,
                    int64_t traj_offset, int32_t *out, void *stream);
/* One basis-state sample per trajectory drawn from |psi'|^2 in index order (stream 1, index 0),
 * with the frame's x mask applied to the result: out[n_traj] (uint64).  For wide registers
 * (dtc_qasm.py measure-all).  scratch: n_traj * 2^(n_local-12 or 0) doubles. */
int dtc_sample_states(const void *state, int n_local, int64_t n_traj, uint64_t seed, int64_t traj_offset,
                      const uint64_t *fx_or_null, double *scratch, uint64_t *out, void *stream);
/* n_samples basis-state samples per trajectory (sample s uses Philox index s): out[n_traj][n_samples].  Shots of an
 * ideal circuit that measures more than 12 qubits (dtc_qasm.py measure-all at L = 20). */
int dtc_sample_states_multi(const void *state, int n_local, int64_t n_traj, int n_samples, uint64_t seed,
                            int64_t traj_offset, const uint64_t *fx_or_null, double *scratch, uint64_t *out,
                            void *stream);

/* ---- density-matrix primitives (exact noisy evolution for small n; Aer method density_matrix).
 *      rho is a 2n-qubit vector: index = row + 2^n * col.  n <= 13. ------------------------- */
int dtc_dm_init(void *rho, int n, uint64_t basis_index, void *stream);
int dtc_dm_rot(void *rho, int n, int qubit, double theta, void *stream);            /* RX(theta) rho RX^dag */
int dtc_dm_diag(void *rho, int n, int n1, const int32_t *q1, const double *a,
                int n2, const int32_t *qi, const int32_t *qj, const double *b, void *stream);
int dtc_dm_pauli_channel(void *rho, int n, int qubit, double px, double py, double pz, void *stream);
/* General single-qubit channel (thermal relaxation, amplitude damping, Kraus sets, reset: the non-Pauli part of a
 * device-calibrated noise model, NoiseModel.from_backend, fast.py:77-78): superop = 4 x 4 complex matrix, row major, (re, im)
 * pairs (32 doubles, host), acting on the (row bit, column bit) block of `qubit` with block index = row + 2 col:
 * out[i] = sum_j superop[i][j] in[j].  One sweep of rho. */
int dtc_dm_superop(void *rho, int n, int qubit, const double *superop, void *stream);
int dtc_dm_probs(const void *rho, int n, int k, const int32_t *qubits, double *out, void *stream);
/* A whole density-matrix program in one call (Aer method density_matrix inside run(), fast.py:211, for n <= 13).  Segments in
 * circuit order: seg_type[i] = 0 rotations RX(val) on q0, 1 diagonal terms exp(-i val Z_q0 / 2) (q1 < 0) or
 * exp(-i val Z_q0 Z_q1 / 2), 2 Pauli channels (probs[k][3] = pX, pY, pZ) on q0; items of segment i are
 * [seg_off[i], seg_off[i+1]).  A rotation segment and the channel segment that follows it are applied per qubit in one
 * sweep of rho for up to six qubits at a time; a diagonal segment is folded into the load of the next sweep.
 * n_sweeps_or_null: number of passes over rho the call made. */
int dtc_dm_run(void *rho, int n, int n_seg, const int32_t *seg_type, const int32_t *seg_off, const int32_t *q0,
               const int32_t *q1, const double *val, const double *probs, int *n_sweeps_or_null, void *stream);

/* ---- sharded statevector support (top log2(P) qubits global) ------------------------------ */
/* Pack / unpack for the all-to-all that exchanges the g = log2(P) global qubits with local qubits
 * lq[0..g-1]: send chunk d (destined to rank d) = amplitudes whose lq bits spell d.
 * pack: out[d][j] = state[insert bits]; unpack is the inverse after the exchange. */
int dtc_shard_pack(const void *state, void *out, int n_local, int g, const int32_t *lq, void *stream);
int dtc_shard_unpack(const void *in, void *state, int n_local, int g, const int32_t *lq, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DTCSIM_H */
