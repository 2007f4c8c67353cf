/* b200fdtd.h -- C ABI of the B200-native FDTD engine (libb200fdtd.so).
 *
 * This is the drop-in boundary for the ONE hot call under spinsphotonics/pjz:
 *
 *     fields = fdtdz_jax.fdtdz(epsilon=, dt=, source_field=, source_waveform=,
 *                              source_position=, absorption_mask=, pml_kappa=, pml_sigma=,
 *                              pml_alpha=, pml_widths=, output_steps=,
 *                              use_reduced_precision=, launch_params=, offset=)
 *                                            -- /root/reference/src/pjz/_field.py:254-269
 *
 * whose implementation is the un-vendored PyPI package fdtdz>=1.1.3
 * (/root/reference/setup.py:26).  In the reference that python function binds a JAX primitive
 * lowered to an XLA GPU custom call `void fn(cudaStream_t, void** buffers, const char* opaque,
 * size_t opaque_len)`.  The entry points below are what that FFI binds instead:
 *
 *   b200fdtd_run            <- the custom call's body (device buffers, caller's stream)
 *   b200fdtd_xla_custom_call<- the legacy XLA custom-call symbol itself (same ABI as fdtd-z's)
 *   b200fdtd_run_host       <- convenience for hosts without a device-array type (NumPy)
 *   b200fdtd_workspace_bytes / b200fdtd_output_bytes / b200fdtd_num_outputs
 *                           <- shape/scratch inference the python wrapper needs
 *
 * Conventions: plain pointers and sizes only, no C++/torch/JAX types; every function returns
 * 0 on success or a B200FDTD_E* code (never throws, never aborts); the message of the last
 * failure on the calling thread is b200fdtd_last_error().  Buffers are owned by the caller.
 * b200fdtd_run is asynchronous w.r.t. the host and ordered on the stream it is given; it is
 * re-entrant per (device, stream) and keeps no global mutable state.
 *
 * All arrays are C-order (last index fastest, i.e. z fastest), float32:
 *   epsilon          (3, xx, yy, zz)      permittivity of the sub-volume at `off_*`; the rest of
 *                                          the (X, Y, Z) domain is edge-replicated
 *   source_field     (2, 1, Y, Z) | (2, X, 1, Z) | (2, 2, X, Y, 1)     (source_axis 0 | 1 | 2)
 *   source_waveform  (tt, 2)
 *   absorption_mask  (3, X, Y)
 *   pml_kappa/sigma/alpha (Z, 2)           column 0: integer-z nodes, column 1: z+1/2 nodes
 *   output           (n_out, 3, xx, yy, zz) E after step n for n in range(out_start,out_stop,out_step)
 */
#ifndef B200FDTD_H_
#define B200FDTD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200FDTD_ABI_VERSION 1

enum {
  B200FDTD_OK = 0,
  B200FDTD_EINVAL = 1,     /* bad descriptor / shapes / null pointer          */
  B200FDTD_EWORKSPACE = 2, /* workspace too small                             */
  B200FDTD_ECUDA = 3,      /* a CUDA call failed (see b200fdtd_last_error)    */
  B200FDTD_EUNSUPPORTED = 4
};

/* Index of each device buffer in `inputs[]`. */
enum {
  B200FDTD_IN_EPSILON = 0,
  B200FDTD_IN_SOURCE_FIELD = 1,
  B200FDTD_IN_SOURCE_WAVEFORM = 2,
  B200FDTD_IN_ABSORPTION_MASK = 3,
  B200FDTD_IN_PML_KAPPA = 4,
  B200FDTD_IN_PML_SIGMA = 5,
  B200FDTD_IN_PML_ALPHA = 6,
  B200FDTD_NUM_INPUTS = 7,      /* mandatory inputs                                          */
  B200FDTD_IN_PROJECTION = 7,   /* optional 8th input, read only when desc->proj_rows > 0     */
  B200FDTD_MAX_INPUTS = 8
};

/* Kernel selection (launch_params). */
enum {
  B200FDTD_KERNEL_AUTO = 0,
  B200FDTD_KERNEL_TWOPASS = 1,  /* one H launch + one E launch per step, in place          */
  B200FDTD_KERNEL_SYSTOLIC = 2, /* one persistent launch: fused H+E x-sweep, L2-pipelined
                                   stages, operands staged through registers               */
  B200FDTD_KERNEL_SYSTOLIC_ASYNC = 3, /* same protocol; operands staged through a cp.async
                                   shared-memory ring, service warp for the protocol       */
  B200FDTD_KERNEL_RESERVED4 = 4, /* (was a TMA-staged variant: measured 17-28 Gcell/s against 76,
                                   removed in round 2; the value is rejected)               */
  B200FDTD_KERNEL_SYSTOLIC_LEAN = 5 /* same protocol; one warp per column pair, no CTA barrier in
                                   the plane loop (fp32, z-column of exactly 32 vectors)     */
};

/* Static description of one engine call (everything that is a python scalar/tuple at
 * /root/reference/src/pjz/_field.py:254-269 plus the array shapes). */
typedef struct b200fdtd_desc {
  uint32_t struct_bytes;        /* = sizeof(b200fdtd_desc)                                  */
  uint32_t abi_version;         /* = B200FDTD_ABI_VERSION                                   */
  int32_t X, Y, Z;              /* full domain: absorption_mask.shape[1:], pml_*.shape[0]   */
  int32_t xx, yy, zz;           /* epsilon.shape[1:]                                        */
  int32_t off_x, off_y, off_z;  /* `offset`                                                 */
  int32_t tt;                   /* source_waveform.shape[0] = number of time steps          */
  int32_t source_axis;          /* 0 | 1 | 2, from source_field's shape                     */
  int32_t source_position;      /* `source_position`                                        */
  int32_t pml_lo, pml_hi;       /* `pml_widths`                                             */
  int32_t out_start, out_stop, out_step; /* `output_steps`                                  */
  int32_t use_reduced_precision;/* 1: E, H and dt/epsilon held as fp16, fp32 arithmetic     */
  float dt;                     /* `dt`                                                     */
  /* `launch_params` (all 0 = choose automatically) */
  int32_t kernel;               /* B200FDTD_KERNEL_*                                        */
  int32_t tile_y;               /* systolic: y-columns owned per CTA                        */
  int32_t stages;               /* systolic: time steps in flight along the x sweep         */
  int32_t threads;              /* CTA size override                                        */
  int32_t prefetch;             /* systolic_async: planes of prefetch distance (1..3)       */
  int32_t cols;                 /* systolic_async: columns per compute thread (1 | 2)       */
  /* Fused frequency projection (replaces the snapshot dump + pinv einsum of
   * /root/reference/src/pjz/_field.py:272-279).  proj_rows = R > 0: inputs[7] is a (R, n_out)
   * float32 matrix W and the output is (R, 3, xx, yy, zz) with
   *   out[r] = sum_s W[r][s] * snapshot_s   (accumulated in snapshot order with one fmaf each),
   * formed inside the time-stepping kernels; no snapshot is ever written.  0: snapshots. */
  int32_t proj_rows;
  int32_t reserved;
} b200fdtd_desc;

/* ABI version of the loaded library. */
int b200fdtd_abi_version(void);

/* Message for the last non-zero return on this thread ("" if none). */
const char* b200fdtd_last_error(void);

/* Validates `desc`; returns B200FDTD_OK or an error code. */
int b200fdtd_validate(const b200fdtd_desc* desc);

/* Number of snapshots len(range(out_start, out_stop, out_step)); < 0 on invalid desc. */
int b200fdtd_num_outputs(const b200fdtd_desc* desc);

/* Bytes of the output array, float32 (n_out, 3, xx, yy, zz) -- or (proj_rows, 3, xx, yy, zz)
 * with the fused projection; 0 on invalid desc. */
size_t b200fdtd_output_bytes(const b200fdtd_desc* desc);

/* Scratch bytes b200fdtd_run needs in device memory; 0 on invalid desc. */
size_t b200fdtd_workspace_bytes(const b200fdtd_desc* desc);

/* The engine call.  `inputs[B200FDTD_NUM_INPUTS]` and `outputs[1]` are DEVICE pointers on the
 * current device; `workspace` is device scratch of at least b200fdtd_workspace_bytes(desc)
 * (256-byte aligned), or NULL to let the engine allocate/free it stream-ordered
 * (cudaMallocAsync).  `stream` is a cudaStream_t passed as void*.  Returns immediately after
 * enqueueing; errors raised later by the device surface at the caller's next sync. */
int b200fdtd_run(const b200fdtd_desc* desc, const void* const* inputs, void* const* outputs,
                 void* workspace, size_t workspace_bytes, void* stream);

/* Same call with HOST buffers: allocates device memory on `device`, copies the inputs up,
 * runs, copies the snapshots back, synchronises and frees.  For NumPy-style callers. */
int b200fdtd_run_host(const b200fdtd_desc* desc, const void* const* host_inputs,
                      void* const* host_outputs, int device);

/* Legacy XLA GPU custom-call entry (the calling convention of fdtd-z's own custom call):
 * buffers[0..6] = inputs in the order above, buffers[7] = output, buffers[8] = workspace
 * (declared by the python wrapper as a second, scratch result); `opaque` = the bytes of a
 * b200fdtd_desc.  Failures are reported through b200fdtd_last_error() and leave the output
 * untouched (XLA's legacy API has no status channel: prefer the _status entry below). */
void b200fdtd_xla_custom_call(void* stream, void** buffers, const char* opaque,
                              size_t opaque_len);

/* The same call under XLA's status-returning convention (API_VERSION_STATUS_RETURNING,
 * xla/service/custom_call_status.h): on failure the message of b200fdtd_last_error() is handed to
 * XLA through XlaCustomCallStatusSetFailure (resolved from the host process at run time), so a bad
 * descriptor or an undersized scratch buffer raises in Python instead of yielding an unwritten
 * result.  `opaque` may carry, after the descriptor, a uint64 with the scratch bytes the wrapper
 * declared at trace time; the run is refused when this device's plan needs more.  This is the
 * entry INTEGRATION.md registers. */
typedef struct b200fdtd_xla_status b200fdtd_xla_status;   /* = XlaCustomCallStatus */
void b200fdtd_xla_custom_call_status(void* stream, void** buffers, const char* opaque,
                                     size_t opaque_len, b200fdtd_xla_status* status);

/* b200fdtd_run_host allocates from a private stream-ordered pool (one per device) that keeps its
 * blocks between calls; this returns them to the driver.  The device's default pool is never
 * touched. */
int b200fdtd_host_pool_trim(int device);

/* Introspection for tests/benchmarks: fills `info[8]` with what the AUTO policy would run for
 * `desc` on the current device: {kernel, tile_y, stages, threads, ctas, smem_bytes,
 * launches_per_run, l2_window_bytes>>20}.  */
int b200fdtd_plan_info(const b200fdtd_desc* desc, int64_t* info);


/* ---- Adjoint product-reduce -----------------------------------------------------------------------
 * The gradient of pjz.scatter's custom_vjp in one pass (replaces the N^2 full-volume temporaries
 * grads[i][j] = F_i F_j / a_i of /root/reference/src/pjz/_field.py:380-382 and their reduction
 * against the cotangents, :393-398):
 *     out[v] = sum_{i,j < nports} sum_{w < ww} Re( coef[i][j][w] * F_i[w][v] * F_j[w][v] )
 * fields[i] : device pointer to port i's phasor field, complex64 (ww, nvox), nvox = 3*xx*yy*zz
 * coef      : device complex64 (nports, nports, ww)  (= conj(cotangent_ij) / amplitude_i)
 * out       : device float32 (nvox).   nports <= 16, ww <= 64.  Asynchronous on `stream`. */
int b200fdtd_adjoint_reduce(int nports, int ww, size_t nvox, const void* const* fields,
                            const void* coef, void* out, void* stream);


/* ---- Snapshot projection and port overlaps ---------------------------------------------------------
 * b200fdtd_project: the phasor extraction of /root/reference/src/pjz/_field.py:272-279
 * (`einsum("ij,j...->i...", pinv(phases.T), fields)`, then `outputs[:ww] + 1j * outputs[ww:]`) in
 * one pass over the snapshots.  snapshots (n_out, nvox) float32 = the engine's output; weights
 * (2*ww, n_out) float32 row-major = the pseudo-inverse; out (ww, nvox) complex64.  Device pointers,
 * asynchronous on `stream`.
 * b200fdtd_overlaps: the two-plane mode overlaps of :305-338 for every (field, port) pair at once:
 *   vals[f][m][k][w] = sum_{c<2, u, v} modes[m][w][c][u][v] * fields[f][w][comp_c][plane k of port m]
 * fields[f] (ww, 3, xx, yy, zz) complex64; modes[m] (ww, 2, U, V) complex64 on the plane normal to
 * axes[m] (the two transverse components in ascending order); planes (nports, 2) int32 HOST array;
 * vals (nfields, nports, 2, ww) complex64 device.  nfields, nports <= 16. */
int b200fdtd_project(int ww, int n_out, size_t nvox, const void* snapshots, const void* weights,
                     void* out, void* stream);
int b200fdtd_overlaps(int nfields, int nports, int ww, int xx, int yy, int zz,
                      const void* const* fields, const void* const* modes, const int* axes,
                      const int* planes, void* vals, void* stream);


/* ---- Waveguide-mode operator --------------------------------------------------------------------
 * y = op(x): the shifted waveguide operator of /root/reference/src/pjz/_mode.py:22-51 applied to
 * all ww frequencies and mm trial vectors in one launch (the body of pjz.mode's subspace
 * iteration, :101-145; the iteration itself is driven by pjz_b200/_mode_gpu.py).
 * eps (3,uu,vv) in "propagate-along-z" form, omega (ww), shift (ww), x and y (ww,2,uu,vv,mm):
 * all device float32.  Asynchronous on `stream`. */
int b200fdtd_mode_operator(int ww, int uu, int vv, int mm, const void* eps, const void* omega,
                           const void* shift, const void* x, void* y, void* stream);


/* ---- Permittivity renderer --------------------------------------------------------------------------
 * pjz.render / pjz.epsilon (/root/reference/src/pjz/_epsilon.py:10-155) on the GPU, writing
 * epsilon straight into the engine's input layout.  layers (ll, 2m*xx, 2m*yy), layer_pos (ll-1),
 * grid_start / grid_end (zz, 2) [column 0: Ex/Ey cells, column 1: Ez cells], out (3, xx, yy, zz):
 * all device float32; workspace: b200fdtd_render_workspace_bytes(ll, xx, yy, zz) device bytes
 * (256-byte aligned).
 * Asynchronous on `stream`. */
size_t b200fdtd_render_workspace_bytes(int ll, int xx, int yy, int zz);
int b200fdtd_render(int ll, int xx, int yy, int zz, int m, const void* layers,
                    const void* layer_pos, const void* grid_start, const void* grid_end,
                    int use_simple_averaging, void* workspace, void* out, void* stream);

/* Vector-Jacobian product of b200fdtd_render (the reference differentiates pjz.render with
 * jax.grad, /root/reference/tests/test_layers.py:179-188).  grad_out (3, xx, yy, zz) float32 =
 * d loss / d epsilon; results: grad_layers (ll, 2m*xx, 2m*yy) float32 = d loss / d layers, and
 * grad_overlap (2, 2, ll, zz) float64 = d loss / d [u | u*z][column][layer][z], the gradient with
 * respect to the layer/cell overlap table of /root/reference/src/pjz/_epsilon.py:47-66 from which
 * the caller forms d loss / d layer_pos (pjz_b200/_epsilon.py does it with four tensor ops).
 * workspace: b200fdtd_render_backward_workspace_bytes(...) device bytes, 256-byte aligned.
 * All pointers are device pointers; asynchronous on `stream`. */
size_t b200fdtd_render_backward_workspace_bytes(int ll, int xx, int yy, int zz);
int b200fdtd_render_backward(int ll, int xx, int yy, int zz, int m, const void* layers,
                             const void* layer_pos, const void* grid_start, const void* grid_end,
                             int use_simple_averaging, const void* grad_out, void* workspace,
                             void* grad_layers, void* grad_overlap, void* stream);


/* ---- Stepping sessions ---------------------------------------------------------------------------
 * For callers that drive the time loop themselves -- the x-slab domain decomposition of
 * pjz_b200/_decomp.py exchanges halo planes between GPUs after every half-step.  A session
 * binds a descriptor, the input arrays and a caller-owned workspace (no internal allocation),
 * prepares the coefficients, and then advances one half-step per call with the per-step
 * kernels, in place (or whole steps with b200fdtd_session_advance, below).  Between calls the caller may read and write field planes directly in the
 * workspace (b200fdtd_session_layout says where they are).  Everything is ordered on the
 * stream passed to each call.  There is no reference counterpart (fdtd-z is single-GPU). */
typedef struct b200fdtd_session b200fdtd_session;

size_t b200fdtd_session_workspace_bytes(const b200fdtd_desc* desc);

int b200fdtd_session_create(const b200fdtd_desc* desc, const void* const* inputs,
                            void* const* outputs, void* workspace, size_t workspace_bytes,
                            void* stream, b200fdtd_session** session);

/* H^{n+1/2} <- H^{n-1/2}, E^n  (all X planes of the session's domain). */
int b200fdtd_session_step_h(b200fdtd_session* session, void* stream);

/* E^{n+1} <- E^n, H^{n+1/2}; adds the source of step n; writes the snapshot if n is an output step. */
int b200fdtd_session_step_e(b200fdtd_session* session, int n, void* stream);

/* info[8] = {byte offset of Ex, byte offset of Hx, bytes between components, bytes per x-plane,
 *            padded z extent Zp, bytes per element, X, Y}: component c of E lives at
 *            workspace + info[0] + c*info[2], laid out [X][Y][Zp]. */
int b200fdtd_session_layout(const b200fdtd_session* session, int64_t* info);

/* Whole steps [n0, n0+nsteps) in one call.  A session whose descriptor selects
 * B200FDTD_KERNEL_AUTO / _SYSTOLIC_LEAN on a geometry that kernel supports runs them as ONE
 * persistent launch of the systolic kernel (the y-slab decomposition with ghost zones of
 * pjz_b200/_decomp.py advances G steps between halo exchanges); its state is ping-ponged: the
 * fields and psiH after n steps live in buffer set n & 1 (psiE is updated in place).  Other
 * sessions loop over step_h / step_e.  step_h / step_e are refused on a systolic session. */
int b200fdtd_session_advance(b200fdtd_session* session, int n0, int nsteps, void* stream);

/* info[16]: entries 0..7 as b200fdtd_session_layout, then
 *   [8]  byte offset of Ex of buffer set 1 (-1: per-step session, single set)
 *   [9]  byte offset of psiH[0] of set 0; psiH[1], psiE[0], psiE[1] follow at multiples of [10]
 *   [10] bytes per psi array, laid out [X][Y][info[12]] floats (only the PML z-groups)
 *   [11] byte offset of psiH[0] of set 1 (-1: single set)      [12] psi floats per column
 *   [13] 1 if the state is ping-ponged (systolic session)      [14] kernel   [15] stages */
int b200fdtd_session_layout2(const b200fdtd_session* session, int64_t* info);

/* ---- y-slab sessions: halo exchange from inside the kernel over peer-mapped memory ---------------
 * Domain decomposition along y for one process per GPU (pjz_b200/_decomp.py:P2PSlabRun; there is
 * no reference counterpart).  The local domain of `desc` has Y = owned columns + 2: the tiles of
 * the persistent kernel cover the owned columns [ylo, yhi) = [1, Y-1) and the columns 0 and Y-1 are
 * ghosts that the NEIGHBOURING GPUs fill: a courier CTA of every launch copies each newly
 * finished plane of the slab's edge columns into the neighbour's ghost column (peer-mapped stores
 * over NVLink) and then the edge tile's progress counter into the neighbour's mirror slot
 * (st.release.sys), so the neighbour's edge tiles depend on them exactly as on a local tile.  One
 * launch advances any number of steps; nothing is exchanged by the host.  Requirements: fp32 storage, 125 <= Z <= 128
 * (the warp-per-column-pair kernel); every rank uses the same local shape and launch parameters;
 * every workspace is mapped into its two neighbours' address space (CUDA IPC or a VMM export --
 * the caller's business; pass the local base for a neighbour that is this rank itself).
 *   1. b200fdtd_session_workspace_bytes_slab / _create_slab   (as the plain versions, + the range)
 *   2. exchange workspace addresses; b200fdtd_session_set_peers(session, low, high)
 *   3. per launch: barrier; b200fdtd_session_slab_reset; barrier; b200fdtd_session_advance.
 *      (The first barrier says every rank has finished its previous launch -- a neighbour still
 *      running would see its mirror slots cleared, or its ghosts overwritten, under its feet.)
 *      b200fdtd_session_advance returns B200FDTD_EINVAL when no reset ran since the last launch:
 *      stale counters would make every dependency look met. */
size_t b200fdtd_session_workspace_bytes_slab(const b200fdtd_desc* desc, int ylo, int yhi);
int b200fdtd_session_create_slab(const b200fdtd_desc* desc, const void* const* inputs,
                                 void* const* outputs, void* workspace, size_t workspace_bytes,
                                 void* stream, int ylo, int yhi, b200fdtd_session** session);
int b200fdtd_session_set_peers(b200fdtd_session* session, void* workspace_low_neighbour,
                               void* workspace_high_neighbour);
int b200fdtd_session_slab_reset(b200fdtd_session* session, void* stream);

/* Peer-mappable device memory for slab workspaces: cudaMalloc + CUDA IPC.  _export writes the
 * 64-byte IPC handle of an allocation made by _alloc; _open maps another process's handle on the
 * current device (enabling peer access); _close unmaps; _free releases. */
int b200fdtd_peer_alloc(size_t bytes, void** ptr);
int b200fdtd_peer_free(void* ptr);
int b200fdtd_peer_export(void* ptr, unsigned char handle[64]);
int b200fdtd_peer_open(const unsigned char handle[64], void** ptr);
int b200fdtd_peer_close(void* ptr);

void b200fdtd_session_destroy(b200fdtd_session* session);

#ifdef __cplusplus
}
#endif
#endif /* B200FDTD_H_ */
