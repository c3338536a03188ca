/* tppvof.h — C-ABI of libtppvof.so, the B200-native drop-in for the OpenFOAM-13
 * `incompressibleVoF` PIMPLE time step of elvis-aguero/openfoam-TPP.
 *
 * Boundary being replaced.  The reference has no in-process FFI for this path: its Python
 * orchestrator shells out to OpenFOAM,
 *     run_case_local -> subprocess.run(["make","-C",case,"run"|"resume",...], check=True)
 *                                              (/root/reference/main.py:333-348)
 *     make run/resume -> [mpirun -np N] foamRun [-parallel]
 *                                              (circularSloshingTank/Makefile:71-99)
 * so the entry points below are what a ctypes binding of that call needs (INTEGRATION.md
 * shows the stub): open a solver on a mesh + dictionaries (tpp_create), advance it the way
 * `foamRun` does (tpp_run_to_write / tpp_step), read fields back for the time-directory
 * writer (tpp_get), and the per-step probes log of system/functions:17-33.
 *
 * Conventions: plain C, caller owns every buffer it passes, the handle owns all device and
 * host memory it allocates.  Every function returns 0 (or a count) on success and a
 * negative code on failure; tpp_last_error() gives the message.  No exceptions cross the
 * ABI.  There is no CPU path: tpp_create fails when no CUDA device is usable.
 * Handles are independent (one CUDA stream each) and may be driven from different host
 * threads; one handle is not re-entrant.
 */
#ifndef TPPVOF_H
#define TPPVOF_H

#ifdef __cplusplus
extern "C" {
#endif

/* boundary conditions of 0/U, 0/alpha.water, 0/p_rgh (circularSloshingTank/0/*:22-31) */
enum { TPP_U_MOVING_WALL = 0, TPP_U_PRESSURE_INLET_OUTLET = 1 };
enum { TPP_A_ZERO_GRADIENT = 0, TPP_A_INLET_OUTLET = 1 };
enum { TPP_P_FIXED_FLUX = 0, TPP_P_TOTAL_PRESSURE = 1 };
/* a patch whose three codes are -1 is a processor patch (multi-GPU halo) */

/* constant/polyMesh as gmshToFoam leaves it (Makefile:73): OpenFOAM face/cell order */
typedef struct {
    int n_points, n_faces, n_internal, n_cells, n_patches;
    const double* points;    /* n_points x 3, undisplaced */
    const int* face_offsets; /* n_faces + 1 */
    const int* face_labels;
    const int* owner;        /* n_faces */
    const int* neighbour;    /* n_internal */
    const int* patch_start;  /* n_patches */
    const int* patch_size;
    const int* patch_bc_u;
    const int* patch_bc_alpha;
    const int* patch_bc_p;
    const double* patch_inlet_alpha; /* inletOutlet inletValue (0/alpha.water:27) */
    const double* patch_p0;          /* totalPressure p0 (0/p_rgh:30) */
    const int* patch_neighb_proc;    /* processor patches: neighbour rank (boundary:neighbProcNo); may be NULL */
} tpp_mesh_t;

/* one entry of system/fvSolution:solvers (p_rgh:42-48, p_rghFinal:50-66) */
typedef struct {
    int type;      /* 0 = PCG, 1 = GAMG */
    int precond;   /* PCG: 0 = DIC (run as Jacobi-PCG on the GPU), 1 = GAMG */
    int smoother;  /* requested OpenFOAM smoother (0 DIC, 1 DICGaussSeidel, 2 GaussSeidel);
                      sequential sweeps have no parallel form: see DESIGN.md */
    double tolerance, rel_tol;
    int max_iter;
    int n_vcycles, n_pre_sweeps, n_post_sweeps, n_finest_sweeps;
    int n_cells_coarsest, merge_levels;
} tpp_solver_t;

typedef struct {
    /* system/controlDict:19-51 */
    double start_time, end_time, delta_t, write_interval;
    double max_co, max_alpha_co, max_delta_t;
    int adjust_time_step;
    /* constant/g:18, physicalProperties.{water,air}:17-21, phaseProperties:19 */
    double g[3];
    double rho1, rho2, nu1, nu2;
    double sigma; /* 0 in every reference case; > 0: continuum surface force, zeroGradient walls (no contact angle) */
    /* system/fvSolution:19-23 (+ MULES nLimiterIter), system/fvSchemes:30 (cAlpha) */
    int n_alpha_subcycles, n_alpha_corr, n_limiter_iter;
    double c_alpha;
    /* system/fvSolution:78-87 */
    int n_correctors, n_non_orth;
    double p_ref_point[3], p_ref_value;
    tpp_solver_t p_rgh, p_rgh_final;
    /* constant/dynamicMeshDict:17-44 + constant/6DoF.dat (generate_motion.py:13-42) */
    double cofg[3];
    int n_motion;         /* 0: static mesh */
    const double* motion; /* n_motion x 7: t, tx ty tz, rx ry rz (degrees) */
} tpp_config_t;

typedef struct tpp_solver* tpp_handle;

/* Build a solver on CUDA device `device` (>= 0).  Fields start at zero / rho2; set the
 * start fields with tpp_set, then call tpp_init_fields once. */
int tpp_create(const tpp_mesh_t* mesh, const tpp_config_t* cfg, int device, tpp_handle* out);
int tpp_destroy(tpp_handle);
const char* tpp_last_error(void);
const char* tpp_version(void);

/* Named arrays (host buffers, FP64).  Names: alpha U p_rgh p rho phi Uf (restart state),
 * alpha_b U_b p_rgh_b rho_b (boundary values), points (moved mesh points), V C Sf ...
 * tpp_size returns the element count, tpp_get/tpp_set copy device<->host. */
long tpp_size(tpp_handle, const char* name);
long tpp_get(tpp_handle, const char* name, double* out, long cap);
long tpp_set(tpp_handle, const char* name, const double* in, long n);
/* device pointer of a named array for zero-copy glue (torch.from_blob / DLPack);
 * borrowed, valid until the next tpp_step / tpp_run_to_write / tpp_destroy */
int tpp_device_ptr(tpp_handle, const char* name, void** ptr, long* n);

/* boundary values and mixture density from the fields just set (start / restart) */
int tpp_init_fields(tpp_handle);
/* restart: deltaT of the run being resumed (<time>/uniform/time) */
int tpp_set_delta_t(tpp_handle, double delta_t);
/* place the solver at time t with step deltaT (restart / tests): re-evaluates the motion */
int tpp_set_time(tpp_handle, double t, double delta_t);

/* n full time steps: Courant -> deltaT -> mesh motion -> alphaPredictor -> momentum
 * assembly -> pressure correctors  (what one `foamRun` iteration does) */
int tpp_step(tpp_handle, int n);
/* step until a write time or end_time: 1 = write time reached, 0 = end of run,
 * 2 = max_steps exhausted */
int tpp_run_to_write(tpp_handle, long max_steps);
/* one named stage on the current state (parity tests): courant adjustDeltaT advanceTime
 * moveMesh alphaBCs UBCs mixture alphaSubCycle alphaPredictor momentum HbyA pcPrepare
 * pcAssemble pcFinish pcEnd pressureCorrector:0 pressureCorrector:1 */
int tpp_stage(tpp_handle, const char* name);

/* Asynchronous read-back for hosts that stream results out while the solver carries on: a snapshot
 * of the array (OpenFOAM file order) is taken on the solver's stream into a staging buffer owned by the
 * handle, the copy into `out` (pinned host memory, or it will not overlap) runs on a second stream.
 * `out` may be read after tpp_sync().  Requesting the same array again is allowed (its next snapshot
 * waits, on the device, until the previous one has left the staging buffer). */
long tpp_get_async(tpp_handle, const char* name, double* out, long cap);
int tpp_sync(tpp_handle);

/* Integer addressing as the kernels use it (for the bit-exactness tests of SURVEY.md 8a row a1):
 * "owner" [n_faces], "neighbour" [n_internal] (lduAddressing, device face order), the cell -> face
 * ELL table "cf" / "cn" [W x nCp, slot-major: (face << 1) | isNeighbourSide, other cell or -1], and
 * "layout" = {nC, nCp, W, nI, nB, nGhost}.  Returns the array length. */
long tpp_get_int(tpp_handle, const char* name, int* out, long cap);

/* out[16]: t, deltaT, step, Co, alphaCo, iters/initial/final residual of the last p_rgh and
 * p_rghFinal solves, reference cell, deltaN, write index, AMG levels, kernel launches */
int tpp_info(tpp_handle, double* out16);

/* The reference's interface metric, computed on the device from the current alpha.water
 * (extract_interface, main.py:727-806, without PyVista): cell values averaged to the mesh points,
 * the `iso` contour (0.5) = one point per mesh edge whose end values straddle it, on the mesh at its
 * current rigid position.  out5 = max z, min z, mean z, number of contour points (the four columns of
 * postProcessing/interface/interface_summary.csv) and the time.  Single-rank meshes only. */
int tpp_interface(tpp_handle, double iso, double* out5);

/* Run statistics over the steps since the last reset (what an OpenFOAM log is grepped for:
 * "No Iterations", and the alpha.water volume balance).  out9 (may be NULL): steps, sum and
 * maximum of the p_rgh / p_rghFinal iteration counts [1..4], steps whose p_rghFinal solve stopped at
 * maxIter [5], sum(alpha V) at the reset [6] and now [7], time integral of the alpha flux through
 * the physical boundary [8] ([7] - [6] + [8] is the volume-conservation defect).  Then, reset > 0:
 * start collecting (two small reductions per step); reset == 0: stop; reset < 0: leave as is. */
int tpp_stats(tpp_handle, int reset, double* out9);

/* solve A x = b on the mesh's LDU addressing with the positive Laplacian coefficients
 * `upper` (off-diagonals are -upper); returns iterations */
int tpp_solve(tpp_handle, const tpp_solver_t* ctl, const double* diag, const double* upper,
              const double* b, double* x, double* init_res, double* final_res);

/* `probes` function object: cell labels (or -1 = not found) sampled on p every step */
int tpp_set_probes(tpp_handle, int n, const int* cells);
long tpp_probe_log(tpp_handle, double* out, long cap_rows); /* rows (t, v0..); drains */
int tpp_find_cell(tpp_handle, const double* xyz);

/* multigrid hierarchy: rows / faces of the fine level and of every coarse level; returns the
 * number of levels (fine included) */
int tpp_amg_levels(tpp_handle, int* n_rows, int* n_faces, int cap);
/* out4: levels smoothed kernel by kernel (mesh included; rows distributed over the ranks),
 * levels of the tail (gathered onto every rank, one persistent kernel), rows of the first tail
 * level, CTAs of the tail kernel */
int tpp_amg_layout(tpp_handle, int* out4);

/* run on the caller's CUDA stream (e.g. torch.cuda.current_stream().cuda_stream) instead
 * of the handle's own, so the caller's CUDA events bracket the solver's kernels */
int tpp_use_stream(tpp_handle, void* cuda_stream);
/* per-kernel CUDA-event timing: tpp_profile(h,1) ... steps ... tpp_profile_report -> text
 * lines "kernel launches total_ms"; returns the length, or -(needed) if cap is too small */
int tpp_profile(tpp_handle, int on);
long tpp_profile_report(tpp_handle, char* buf, long cap);

/* Multi-GPU, one case decomposed over ranks (one process per GPU).  The mesh passed to
 * tpp_create is the rank's processorN mesh; its `processor` patches (BC codes -1, listed after
 * the physical patches, faces in the neighbour's matching order as decomposePar writes them)
 * become halo interfaces.  tpp_comm_init joins the ranks over NCCL (the library the host
 * process already loaded: pass its path, e.g. torch's nvidia/nccl/lib/libnccl.so.2), exchanges
 * cudaIpc handles of one peer-memory window per rank through that communicator and finishes the
 * processor-face geometry.  Every stencil kernel is then preceded by a halo exchange - one
 * kernel that stores the owner-side values straight into the neighbours' windows over NVLink
 * and unpacks what arrived (falls back to ncclSend/ncclRecv without peer access) - and every
 * Krylov dot, residual norm and Courant maximum is all-reduced over the same windows (<= 8
 * ranks) or by ncclAllReduce; the multigrid keeps its inter-rank couplings on every level and
 * gathers its smallest levels onto every rank.  The replacement, on one node, of
 * `mpirun -np N foamRun -parallel` (/root/reference/circularSloshingTank/Makefile:78,91). */
int tpp_nccl_unique_id(const char* nccl_path, char* out128);
int tpp_comm_init(tpp_handle, int rank, int n_ranks, const char* id128, const char* nccl_path);
/* the same over two host callbacks (CPU tests with gloo): exchange(user, send[nGhost*ncomp],
 * recv[nGhost*ncomp], ncomp) in ghost order; allreduce(user, vals, n, op) with op 0 sum, 1 max */
int tpp_comm_callbacks(tpp_handle, int rank, int n_ranks,
                       int (*exchange)(void*, const double*, double*, int),
                       int (*allreduce)(void*, double*, int, int), void* user);
/* ghost layout: per processor patch the offset/count in ghost order and the neighbour rank */
int tpp_ghost_layout(tpp_handle, int* n_ghost, int* n_patches, int* off, int* cnt, int* peer, int cap);

/* Case directories inside the library: the boundary as SURVEY.md 8(b) sketches it, for hosts that
 * are not Python and have no FoamFile reader of their own (tools/tpp_foamrun.c is one, in C).  Together
 * these are what `foamRun` does when `make run` / `make resume` start it in a case directory
 * (/root/reference/circularSloshingTank/Makefile:85,98 <- main.py:333-348):
 *   tpp_open       reads constant/polyMesh (ascii or binary), the dictionaries under system/ and constant/
 *                  (incl. dynamicMeshDict and its 6DoF table) and the start fields - `startFrom latestTime`
 *                  (controlDict:19): 0/ after setFields, or the newest complete time directory = resume - and
 *                  builds the solver on `device`.  processor < 0: the whole case, ready to run.  processor = k:
 *                  mesh and fields of `processor<k>/` (decomposePar, Makefile:77), dictionaries from the case
 *                  root; join the ranks with tpp_comm_init / tpp_comm_callbacks, then call tpp_case_start.
 *                  Keywords the solver cannot honour (another scheme, boundary condition, solver ...) are
 *                  errors naming file and keyword (-4), exactly as in the Python host (case.read_config).
 *   tpp_write_time writes the current state as a time directory of the opened case: alpha.water U p_rgh p rho
 *                  phi Uf [polyMesh/points] uniform/time, in controlDict's writeFormat / writePrecision /
 *                  timePrecision.  uniform/time is written last and marks the directory complete.
 *   tpp_run_case   the foamRun loop: steps to each write time (tpp_run_to_write), writes it, appends the
 *                  `probes` rows to postProcessing/probes/<start>/p (single-rank cases; a processor share
 *                  leaves its rows in tpp_probe_log for the host to merge), until endTime or max_steps
 *                  (< 0: no limit).  flags: TPP_RUN_LOG = one "Time = ..." line per write on stdout;
 *                  TPP_RUN_INTERFACE = a row of postProcessing/interface/interface_summary.csv per write time
 *                  (tpp_interface; the file extract_interface makes afterwards, main.py:751-780).  Returns the
 *                  number of steps taken, or a negative code (a diverged / failed run is an error, not an `End`).
 *   tpp_case_query integers of the opened case ("n_cells" "n_faces" "n_internal" "n_points" "n_patches"
 *                  "n_probes" "write_binary") as the return value, names ("start_time" "time" "dir") copied
 *                  into text[cap] with their length returned. */
int tpp_open(const char* case_dir, int processor, int device, tpp_handle* out);
int tpp_case_start(tpp_handle);
int tpp_write_time(tpp_handle);
enum { TPP_RUN_LOG = 1, TPP_RUN_INTERFACE = 2 };
long tpp_run_case(tpp_handle, long max_steps, int flags);
long tpp_case_query(tpp_handle, const char* what, char* text, long cap);
/* internalField of a vol/surface field file (ascii or binary; no handle needed), for hosts that
 * post-process time directories (the reference reads alpha.water for its interface metric,
 * main.py:727-806).  Returns the number of doubles the field holds (values x *n_comp; one value when
 * `uniform`) and copies up to cap of them into out. */
long tpp_read_field(const char* path, double* out, long cap, int* n_comp, int* uniform);

#ifdef __cplusplus
}
#endif
#endif
