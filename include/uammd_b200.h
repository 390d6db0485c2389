/* uammd_b200 - C ABI of the B200-native (sm_100a) engine for UAMMD's two data-parallel hot paths.
 *
 * UAMMD has no FFI: its extension points are C++ template concepts and virtual classes compiled into
 * the user's translation unit. The drop-in therefore has two layers: the C++14 glue
 * headers in include/uammd_b200/ (classes satisfying UAMMD's Interactor / NeighbourList / BDHI-Method
 * concepts) and this C ABI, which is everything those glue classes call. Each entry point cites the
 * reference interface it replaces (file:line relative to the reference's src/).
 *
 * Conventions
 *  - all pointers named d_* are DEVICE pointers owned by the caller (UAMMD's ParticleData), layouts are
 *    the reference's: pos real4 {x,y,z,type}, force real4 {fx,fy,fz,_}, vel real3, energy/virial real.
 *  - every call is asynchronous on the given CUDA stream (passed as void* == cudaStream_t); no hidden
 *    device synchronisation; device scratch is owned by the handle and only (re)allocated when N or the
 *    cell grid grows.
 *  - return value: 0 = ok, <0 = error (UB200_ERR_*); ub200_error_string() describes it. No exceptions
 *    cross the boundary; the glue converts codes into the reference's exceptions.
 *  - precision: the reference fixes `real` per translation unit (-DDOUBLE_PRECISION, global/defines.h:9-11)
 *    so symbols come in _f32 / _f64 flavours.
 */
#ifndef UAMMD_B200_H
#define UAMMD_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define UB200_OK 0
#define UB200_ERR_INVALID_ARGUMENT (-1)
#define UB200_ERR_CUDA (-2)
#define UB200_ERR_ALLOC (-3)
#define UB200_ERR_GRID_TOO_LARGE (-4)
#define UB200_ERR_NOT_BUILT (-5)
#define UB200_ERR_UNSUPPORTED (-6)

const char *ub200_error_string(int code);
/* last cudaError_t seen by the calling thread inside the library (0 if none) */
int ub200_last_cuda_error(void);
/* library version / build info: "uammd_b200 <ver> sm_100a" */
const char *ub200_version(void);

/* ------------------------------------------------------------------------------------------------
 * Path 1a: cell list.  Replaces CellList::update / CellListBase::update
 * (Interactor/NeighbourList/CellList.cuh:145-163, CellList/CellListBase.cuh:124-140) and the
 * ParticleSorter Morton sort beneath it (utils/ParticleSorter.cuh:156-164,243-274).
 * Output arrays are bit-identical to the reference's CellListData (CellListBase.cuh:145-160):
 * sortPos, groupIndex (stable Morton order), cellStart (+VALID_CELL epoch bias), cellEnd.
 * ------------------------------------------------------------------------------------------------ */
typedef struct ub200_celllist ub200_celllist;

typedef struct {
  const uint32_t *d_cellStart; /* [ncells] first sorted index + VALID_CELL; < VALID_CELL means empty */
  const int *d_cellEnd;        /* [ncells] last sorted index + 1 (only meaningful for non-empty cells) */
  const void *d_sortPos;       /* real4[N] positions in sorted order */
  const int *d_groupIndex;     /* int[N] sorted slot -> particle (group) index */
  uint32_t VALID_CELL;
  int cellDim[3];
  int numberParticles;
  /* engine-native view (dense, no epoch trick): bins in Morton-code space */
  const uint32_t *d_binStart;  /* [nbins+1] exclusive prefix of particles per Morton code */
  int nbins;
} ub200_celllist_view;

int ub200_celllist_create(ub200_celllist **out);
int ub200_celllist_destroy(ub200_celllist *cl);

/* cellDim as CellList::createUpdateGrid computes it (CellList.cuh:100-126): int(L/rc), <=3 -> 1 */
int ub200_neighbour_celldim_f32(const float L[3], float rc, int cellDim[3]);

/* d_pos: real4[*]; d_groupIdx: optional int[N] indirection (ParticleGroup::getIndexIterator,
 * ParticleData/ParticleGroup.cuh:304-327) or NULL for identity. periodic[d]==0 marks a non periodic
 * dimension (Box::setPeriodicity utils/Box.cuh:32-39). */
int ub200_celllist_build_f32(ub200_celllist *cl, const void *d_pos, const int *d_groupIdx, int N,
                             const float L[3], const int periodic[3], const int cellDim[3], void *stream);
int ub200_celllist_view_get(ub200_celllist *cl, ub200_celllist_view *view);
/* reads back (synchronising the stream) the NaN / out-of-box flag of the last build
 * (CellList_ns::fillCellList errorFlag, CellListBase.cuh:68-95,258-265). 0 = clean. */
int ub200_celllist_error_flag(ub200_celllist *cl, void *stream, int *flag);

/* ------------------------------------------------------------------------------------------------
 * Path 1b: pair traversal with the LJ transverser. Replaces CellList::transverseList
 * (CellList.cuh:165-182 -> NeighbourList/common.cuh:10-34) specialised for
 * Potential::Radial<LJFunctor>::Transverser (Potential/RadialPotential.cuh:107-127, Potential.cuh:25-83).
 * params: host array [ntypes*ntypes] of {cutOff2, sigma2, epsilonDivSigma2, shift}
 * (LJFunctor::PairParameters, Potential.cuh:31-35). Outputs ACCUMULATE (+=) like Transverser::set; any
 * of d_force / d_energy / d_virial may be NULL (Interactor::Computables, Interactor/Interactor.cuh:94-103).
 * d_globalIdx: optional group->global index map used for the output scatter (NULL = identity).
 * ------------------------------------------------------------------------------------------------ */
int ub200_lj_sum_f32(ub200_celllist *cl, const float *params, int ntypes, void *d_force, float *d_energy,
                     float *d_virial, const int *d_globalIdx, void *stream);

/* same with the parameter table already on the device (Radial<LJFunctor>'s BasicParameterHandler keeps it there,
 * Potential/ParameterHandler.cuh:8-12,62-65): d_params = PairParameters[ntypes*ntypes] */
int ub200_lj_sum_devparams_f32(ub200_celllist *cl, const void *d_params, int ntypes, void *d_force, float *d_energy,
                               float *d_virial, const int *d_globalIdx, void *stream);

/* Multi-GPU particle decomposition: forces only for the home particles whose (group) index lies in
 * [ownerLo, ownerHi); every rank holds all positions and the full cell list (the reference is single-GPU: new
 * functionality). accumulate = 0 writes (x,y,z,0), 1 adds. */
int ub200_lj_sum_owned_f32(ub200_celllist *cl, const float *params, int ntypes, void *d_force, int ownerLo,
                           int ownerHi, int accumulate, void *stream);

/* All-pairs fallback of PairForces: the reference switches to NBody::transverse when the box is no larger than
 * 3 cut-offs in every dimension (Interactor/PairForces.cu:49-53,61-66 -> Interactor/NBodyBase.cuh:46-116). One thread
 * per particle, the others visited in ascending group order and accumulated sequentially like the reference kernel,
 * per-pair minimum image (RadialPotential.cuh:107-127). d_globalIdx: optional group index list (N entries);
 * outputs accumulate (+=). */
int ub200_lj_nbody_f32(const void *d_pos, const int *d_globalIdx, int N, const float L[3], const int periodic[3],
                       const float *params, int ntypes, void *d_force, float *d_energy, float *d_virial, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Path 1b in double precision: PairForces<Potential::LJ, CellList>::sum of a `real = double` build of the reference
 * (-DDOUBLE_PRECISION, global/defines.h; Interactor/PairForces.cu:43-78 with Radial<LJFunctor>::Transverser,
 * Potential/RadialPotential.cuh:107-127, LJFunctor::force / energy Potential/Potential.cuh:37-56, Box::apply_pbc
 * utils/Box.cuh:50-57 - all evaluated in double on the double positions). d_pos: double4[N] (w = type); params: HOST array
 * [ntypes*ntypes] of {cutOff2, sigma2, epsilonDivSigma2, shift} in double; cutOff: the largest cut-off (what
 * Potential::getCutOff returns); d_force double4[N], d_energy / d_virial double[N] accumulate (+=) like Transverser::set,
 * any may be NULL. The neighbour search runs on a single precision copy of the positions with the cut-off padded by the
 * rounding of that copy; boxes <= 3 cut-offs collapse to one cell (CellList.cuh:117-122) and are summed all-pairs with the
 * per-pair minimum image. Results equal the reference's up to the order of the sums.
 * ------------------------------------------------------------------------------------------------ */
typedef struct ub200_lj64 ub200_lj64;
int ub200_lj64_create(ub200_lj64 **out);
int ub200_lj64_destroy(ub200_lj64 *lj);
int ub200_lj_sum_f64(ub200_lj64 *lj, const void *d_pos, int N, const double L[3], const int periodic[3], double cutOff,
                     const double *params, int ntypes, void *d_force, double *d_energy, double *d_virial, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Path 1b': PairForces<Potential::LJ, CellList>::sum in ONE call (Interactor/PairForces.cu:43-78: neighbour-list update +
 * traversal with Radial<LJFunctor>::Transverser, or NBody::transverse for boxes <= 3 cut-offs, :49-53). The forces,
 * energies and virials are the reference's (every pair within the cut-off once, minimum image, self excluded); the
 * neighbour search behind them is the engine's own: particles are binned on a grid of HALF cells (edge >= cutOff / 2,
 * sorted x-fastest, coordinates folded into the box) and every column of half cells is traversed by one warp whose
 * 5 x 5 x (6 + 4) halo is staged into shared memory by the TMA engine (cp.async.bulk + mbarrier). 196 candidates per
 * particle at rho = 0.8, rc = 2.5 instead of the 340 of the 27-cell walk (CellList/NeighbourContainer.cuh:95-138).
 * Grids with a periodic dimension under five half cells fall back to ub200_celllist_build_f32 + ub200_lj_sum_f32.
 * d_pos real4[*]; d_groupIdx optional int[N] indirection (ParticleGroup::getIndexIterator); params as ub200_lj_sum_f32;
 * outputs indexed by d_globalIdx[group index] (NULL = identity). accumulate = 1 adds like Transverser::set,
 * accumulate = 0 writes (fx, fy, fz, 0) (sole interactor: replaces VerletNVE::resetForces + sum; forces only).
 * [ownerLo, ownerHi): only particles whose group index lies in the range are computed and written (0, INT_MAX = all;
 * multi-GPU decompositions, forces only).
 * ------------------------------------------------------------------------------------------------ */
typedef struct ub200_ljengine ub200_ljengine;
int ub200_ljengine_create(ub200_ljengine **out);
int ub200_ljengine_destroy(ub200_ljengine *e);
int ub200_ljengine_sum_f32(ub200_ljengine *e, const void *d_pos, const int *d_groupIdx, int N, const float L[3],
                           const int periodic[3], const float *params, int ntypes, void *d_force, float *d_energy,
                           float *d_virial, const int *d_globalIdx, int accumulate, int ownerLo, int ownerHi, void *stream);
/* The neighbour search alone: builds the half-cell list for this cut-off (UB200_ERR_UNSUPPORTED when a periodic dimension has
 * fewer than five half cells: use ub200_celllist_build_f32 there). ub200_ljengine_view_get exposes it to the header-template
 * traversal of include/uammd_b200/uammd_b200.cuh (b200::ColumnList::transverseList), which runs ANY user Transverser
 * (utils/TransverserUtils.cuh:151-274) over it: d_pos real4[N] in list order, folded into the box and consistent with the
 * cells (w = type); d_index[k] = group index of list slot k; d_cellStart[c] .. d_cellStart[c + 1] = slots of half cell
 * c = x + cells[0] (y + cells[1] z). */
typedef struct {
  const void *d_pos;
  const int *d_index;
  const uint32_t *d_cellStart;
  int cells[3], periodic[3];
  float L[3];
  int numberParticles;
} ub200_ljengine_view;
int ub200_ljengine_build_f32(ub200_ljengine *e, const void *d_pos, const int *d_groupIdx, int N, const float L[3],
                             const int periodic[3], float cutOff, void *stream);
int ub200_ljengine_view_get(ub200_ljengine *e, ub200_ljengine_view *view);
/* the traversal kernel alone over the list of the last ub200_ljengine_sum_f32 call (positions unchanged; kernel timing
 * and profiling). Forces only; UB200_ERR_NOT_BUILT unless that call took the column path. */
int ub200_ljengine_traverse_f32(ub200_ljengine *e, void *d_force, int accumulate, void *stream);
/* path of the last call: 0 = column traversal over the half-cell list, 1 = cell traversal over the reference-layout
 * list, 2 = all pairs (NBody) */
int ub200_ljengine_last_path(ub200_ljengine *e);
/* half cells per dimension of the last column traversal */
int ub200_ljengine_grid(ub200_ljengine *e, int cells[3]);
/* synchronises the stream; 1 = a particle outside a non periodic box / NaN (like CellList_ns::fillCellList's errorFlag,
 * CellListBase.cuh:68-95), 2 = a bulk copy never completed */
int ub200_ljengine_error_flag(ub200_ljengine *e, void *stream, int *flag);

/* Brick domain decomposition of the pair path over the GPUs of one box (the reference is single-GPU: new functionality,
 * SURVEY 8(e); BASELINE config 4 "ghost-cell halo exchange"). Defined on the reference's neighbour grid
 * (CellList::createUpdateGrid, CellList.cuh:100-126; Grid::getCell, utils/Grid.cuh:49-71, same roundings as
 * ub200_celllist_build_f32): rank (kx,ky,kz) of rankGrid owns the cells [floor(k n/p), floor((k+1) n/p)) of every
 * dimension and the particles in them; rank r needs as ghosts the particles of the cells adjacent (27-neighbourhood,
 * periodic wrap like Grid::pbc_cell, utils/Grid.cuh:81-106) to its own. Per particle: d_cell (optional) linear cell
 * index x-fastest, d_owner the owning rank (kx + px (ky + py kz)), d_ghostMask bit r set iff rank r != owner needs the
 * particle as a ghost. At most 32 ranks; every brick must hold at least one cell per dimension. */
int ub200_brick_classify_f32(const void *d_pos, int N, const float L[3], const int periodic[3], const int cellDim[3],
                             const int rankGrid[3], int *d_cell, int *d_owner, uint32_t *d_ghostMask, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU pair path: brick domain decomposition with a ghost-cell halo exchange, one process per GPU of one NVSwitch
 * box (new functionality: the reference is single-GPU; BASELINE config 4). The engine's half-cell grid
 * (ub200_ljengine_*) is cut into rankGrid[0] x rankGrid[1] x rankGrid[2] bricks of whole half cells; a rank owns the
 * particles of its brick and holds, as ghosts, the particles of the two half-cell layers (>= cutOff) around it; its list
 * is built on that window only. The particle state {pos real4, vel real3, global id} lives in the handle, on the device.
 * ONE exchange per step: after the drift every rank stores each particle's 32-byte row straight into the inboxes of its
 * new owner and of every rank that needs it as a ghost (peer-to-peer stores over NVLink, CUDA IPC mappings), raises its
 * flag in every peer, and the receivers append what arrived; particle counts never leave the device and no call below
 * synchronises except ub200_brick_counts. Trajectories are bit-identical to the single-GPU engine's
 * (ub200_md_lj_nve_run_f32) for every rank grid: cells list their particles by global id and the image shifts are the
 * single-GPU ones.
 * Set-up (like ub200_fcm_dist_*): create on every rank, exchange the ub200_comm_ipc_size()-byte blobs of
 * ub200_brick_ipc_export between the ranks (rank order), hand all of them to ub200_brick_ipc_import. Virtual ranks inside
 * one process (tests on one GPU) pass each other's ub200_brick_arena pointers to ub200_brick_attach_local instead and
 * drive the step in two phases (ub200_brick_lj_nve_phase_f32: phase 0 of every rank before phase 1 of any).
 * At most 8 ranks; every brick at least two half cells thick; capacity = 0 picks 1.3 x the mean window population.
 * ------------------------------------------------------------------------------------------------ */
typedef struct ub200_brick ub200_brick;
typedef struct {
  void *d_pos, *d_vel;  /* real4[capacity], real3[capacity]: owned block first, ghosts behind it */
  int *d_gid;           /* global particle ids */
  void *d_force;        /* real4[capacity], meaningful on the owned block */
  int *d_counts;        /* device {nOwned, nLocal} */
  int capacity, rank, world;
  int halfCells[3], window[3], windowOrigin[3];
} ub200_brick_info_t;
int ub200_brick_create(ub200_brick **out, int rank, const int rankGrid[3], const float L[3], const int periodic[3], float cutOff,
                       int numberParticles, int capacity);
int ub200_brick_destroy(ub200_brick *b);
int ub200_comm_ipc_size(void);
int ub200_brick_ipc_export(ub200_brick *b, void *blob);
int ub200_brick_ipc_import(ub200_brick *b, const void *blobsOfAllRanks);
int ub200_brick_arena(ub200_brick *b, void **arena);
int ub200_brick_attach_local(ub200_brick *b, void *const *arenasOfAllRanks);
/* every rank passes the same full arrays (replicated initial condition) and keeps the particles it owns */
int ub200_brick_set_global_state_f32(ub200_brick *b, const void *d_pos, const void *d_vel, int N, void *stream);
/* ownership + ghosts for the current positions: the exchange alone (migration and halo in one pass) */
int ub200_halo_exchange_f32(ub200_brick *b, void *stream);
int ub200_halo_exchange_phase_f32(ub200_brick *b, int phase, void *stream);
/* LJ forces of the owned block over [owned | ghosts] (ub200_ljengine_sum_f32 restricted to the owned particles) */
int ub200_brick_lj_forces_f32(ub200_brick *b, const float *params, int ntypes, void *stream);
/* VerletNVE::forwardTime x nsteps with one PairForces<LJ, CellList> interactor (Integrator/VerletNVE.cu:174-188) on the
 * bricks: kick + drift + exchange, list build over the window, forces of the owned block, kick. */
int ub200_brick_lj_nve_run_f32(ub200_brick *b, const float *params, int ntypes, float dt, int nsteps, void *stream);
int ub200_brick_lj_nve_phase_f32(ub200_brick *b, int phase, const float *params, int ntypes, float dt, int doKick, void *stream);
/* The same loop with one PairForces<Potential::DPD> interactor (BASELINE config 4: "DPD fluid, ghost-cell halo exchange,
 * domain-decomposed over 8 GPUs"). Create the bricks with cutOff = rcut. DPD_impl::ForceTransverser (Potential/DPD.cuh:92-159);
 * sigma = sqrt(2 T) / sqrt(dt) as DPD_impl computes it; ghosts carry their velocities and the pair noise is keyed on the
 * GLOBAL ids (ij = min + N max), so the trajectory is the single-GPU one for every rank grid. The step counter of the noise
 * starts at 1 with the first force evaluation and advances by one per evaluation (DPD.cuh:165). */
int ub200_brick_dpd_nve_run_f32(ub200_brick *b, float A, float gamma, float sigma, float rcut, uint32_t seed, float dt, int nsteps,
                                void *stream);
int ub200_brick_dpd_nve_phase_f32(ub200_brick *b, int phase, float A, float gamma, float sigma, float rcut, uint32_t seed, float dt,
                                  int doKick, void *stream);
int ub200_brick_info(ub200_brick *b, ub200_brick_info_t *info);
/* owned block <-> host buffers (pinned for asynchronous copies): pos real4[n], vel real3[n], ids int[n] (may be NULL);
 * n = the owned count ub200_brick_counts reported. The upload replaces the owned block in place. */
int ub200_brick_download_owned_f32(ub200_brick *b, void *h_pos, void *h_vel, int *h_gid, int n, void *stream);
int ub200_brick_upload_owned_f32(ub200_brick *b, const void *h_pos, const void *h_vel, const int *h_gid, int n, void *stream);
/* diagnostics (environment UB200_BRICK_PROFILE=1 at create time, makes every step synchronous): mean ms per step of
 * push, unpack (including the wait for the peers), list build, traversal, kick */
int ub200_brick_profile(ub200_brick *b, double phases[5]);
/* synchronises the stream: particle counts and the error flag (0 clean; 3 particle outside the window, 4 capacity,
 * 5 inbox overflow, 6 a peer's flag never arrived) */
int ub200_brick_counts(ub200_brick *b, void *stream, int *nOwned, int *nLocal, int *errorFlag);

/* ParticleData::sortParticles (ParticleData/ParticleData.cuh:492-522): d_order[k] = index of the particle that comes k-th in the
 * stable sort by the Morton hash of its cell, cells of L / hashCutOff per dimension (truncated, like hints.hash_box.boxSize /
 * hints.hash_cutOff) - the permutation ParticleSorter::updateOrderByCellHash computes (utils/ParticleSorter.cuh:157-164), bit for
 * bit. scratch: any cell list handle (its contents are replaced). ub200_apply_order = ParticleSorter::applyCurrentOrder
 * (:177-187): d_out[k] = d_in[d_order[k]] for rows of rowBytes bytes (a multiple of 4); in and out must not alias. */
int ub200_particles_sort_order_f32(ub200_celllist *scratch, const void *d_pos, int N, const float L[3], const int periodic[3],
                                   float hashCutOff, int *d_order, void *stream);
int ub200_apply_order(const void *d_in, void *d_out, const int *d_order, int N, int rowBytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Path 3: spectral Ewald Poisson solver for Gaussian charges in a triply periodic box. Replaces
 * Poisson (Interactor/SpectralEwaldPoisson.cuh:84-184, SpectralEwaldPoisson.cu:74-580): same parameter resolution (grid
 * spacing from the tolerance, FFT-friendly grid, Gaussian support, near-field cut-off and table sizes), far field through
 * spread -> FFT -> (-ik, 1) rho / (eps k^2) -> inverse FFT -> interpolation, near field (split > 0) as the reference's three
 * Transversers over a cell list with its two tabulated Green's functions. Precision = the reference's `real` (4 | 8).
 * The reference's Parameters::cells and ::support are accepted but, like there, not used by the constructor.
 * ------------------------------------------------------------------------------------------------ */
typedef struct ub200_poisson ub200_poisson;
typedef struct {
  double L[3];
  double epsilon;     /* permittivity */
  double tolerance;   /* default 1e-5 */
  double gw;          /* Gaussian width of the sources */
  double split;       /* Ewald splitting parameter; <= 0: no splitting (far field only) */
  double upsampling;  /* > 0: grid spacing = 1 / upsampling; <= 0: from the tolerance */
} ub200_poisson_params;
typedef struct {
  int cells[3], support, nTable;
  double h, farFieldGaussianWidth, nearFieldCutOff;
} ub200_poisson_info_t;
int ub200_poisson_create(ub200_poisson **out, int precisionBytes, const ub200_poisson_params *par);
int ub200_poisson_destroy(ub200_poisson *p);
int ub200_poisson_info(ub200_poisson *p, ub200_poisson_info_t *info);
/* Poisson::sum (SpectralEwaldPoisson.cuh:110-122): d_force4 (real4[N], may be NULL) += q E, d_energy (real[N], may be NULL)
 * += q phi; positions real4[N], charges real[N]. */
int ub200_poisson_sum(ub200_poisson *p, const void *d_pos, const void *d_charge, int N, void *d_force4, void *d_energy,
                      void *stream);
/* The reference's exact call pattern: its far field interpolates into the force AND the energy array on every sum() whatever
 * was requested (SpectralEwaldPoisson.cu:561-578), the near field follows the requested computables (:368-410). */
int ub200_poisson_sum_ex(ub200_poisson *p, const void *d_pos, const void *d_charge, int N, void *d_force4, void *d_energy,
                         int nearFieldForce, int nearFieldEnergy, void *stream);
/* Poisson::computeFieldPotentialAtParticles (:124-135): (Ex, Ey, Ez, phi) ADDED to d_fieldPotential4 (real4[N]). */
int ub200_poisson_field_potential(ub200_poisson *p, const void *d_pos, const void *d_charge, int N, void *d_fieldPotential4,
                                  void *stream);

/* ------------------------------------------------------------------------------------------------
 * Path 1d: Verlet (skin) list. Replaces VerletList::update (Interactor/NeighbourList/VerletList.cuh:111-124) and
 * the classes beneath it (VerletList/VerletListBase.cuh:73-199, BasicList/BasicListBase.cuh:76-215): rebuild when
 * a particle moved >= (multiplier - 1) cutOff / 2 since the last rebuild (host-synchronous flag read, like
 * VerletListBase.cuh:191-218). Two lists live behind the handle:
 *  - the ROW LIST of the built-in LJ traversal (ub200_lj_sum_verlet_f32, the fused MD loops): filled from the engine's
 *    half-cell columns, one row per particle, image of each neighbour stored in the entry (lj_vlist.cu). Built at every
 *    rebuild when the grid allows it (>= 5 half cells per periodic dimension, N < 2^27);
 *  - the REFERENCE-LAYOUT list, [k * N + i] over SORTED indices, sortPos refreshed every call, bit-identical to the
 *    reference's VerletListData (BasicListBase.cuh:143-151). Built at every rebuild when the row list does not apply,
 *    otherwise from the first ub200_verletlist_view_get on (which builds it from the positions stored at the last
 *    rebuild; the arrays given to the last update must still be alive) - callers that never read it never pay for it.
 * ------------------------------------------------------------------------------------------------ */
typedef struct ub200_verletlist ub200_verletlist;
typedef struct {
  const int *d_neighbourList;    /* neighbour k of sorted particle i at [k * particleStride + i] (sorted indices) */
  const int *d_numberNeighbours; /* [N] (self included, like the reference) */
  const void *d_sortPos;         /* real4[N] current positions in the sorted order of the last rebuild */
  const int *d_groupIndex;       /* [N] sorted slot -> group index */
  int particleStride, numberParticles, maxNeighboursPerParticle;
  int stepsSinceLastUpdate;      /* VerletList::getNumberOfStepsSinceLastUpdate */
  int rebuilds;                  /* rebuilds since creation */
} ub200_verletlist_view;
int ub200_verletlist_create(ub200_verletlist **out);
int ub200_verletlist_destroy(ub200_verletlist *vl);
int ub200_verletlist_set_cutoff_multiplier(ub200_verletlist *vl, float multiplier); /* VerletList.cuh:169-172, default 1.08 */
/* forceRebuild: the caller's "positions were written / particles were reordered" signal (VerletList.cuh:179-190);
 * the drift check handles everything else. Synchronises the stream when it has to read the drift / overflow flag. */
int ub200_verletlist_update_f32(ub200_verletlist *vl, const void *d_pos, const int *d_groupIdx, int N, const float L[3],
                                const int periodic[3], float cutOff, int forceRebuild, int *rebuilt, void *stream);
int ub200_verletlist_view_get(ub200_verletlist *vl, ub200_verletlist_view *view);
/* VerletList::getNumberOfStepsSinceLastUpdate (VerletListBase.cuh:131) and the rebuild count, without touching the lists */
int ub200_verletlist_stats(ub200_verletlist *vl, int *stepsSinceLastUpdate, int *rebuilds);
/* The row list (tests, tools): neighbour k of the particle in half-cell slot i at d_list[i * stride + k] =
 * slot | image << indexBits, image = (sx + 1) + 3 (sy + 1) + 9 (sz + 1) box lengths to add to the neighbour (13 = none);
 * self excluded. UB200_ERR_NOT_BUILT when the last rebuild made the reference-layout list only. */
typedef struct {
  const int *d_list;
  const int *d_count;            /* [N] */
  const void *d_pos;             /* real4[N] current positions in half-cell order */
  const int *d_index;            /* [N] half-cell slot -> group index */
  int stride, numberParticles, indexBits;
} ub200_verletlist_rows;
int ub200_verletlist_rows_get(ub200_verletlist *vl, ub200_verletlist_rows *rows);
/* LJ transverser over the Verlet list. Replaces VerletList::transverseList (VerletList.cuh:141-159 ->
 * NeighbourList/common.cuh:10-34 with VerletListBase_ns::NeighbourContainer, BasicList/NeighbourContainer.cuh:42-125)
 * for Radial<LJFunctor>; same argument meaning as ub200_lj_sum_f32 (outputs accumulate). */
int ub200_lj_sum_verlet_f32(ub200_verletlist *vl, const float *params, int ntypes, void *d_force, float *d_energy,
                            float *d_virial, const int *d_globalIdx, void *stream);

/* VerletNVE::forwardTime x nsteps over PairForces<LJ, VerletList> (the pair generic_md instantiates): kick+drift, list
 * update (drift check), forces, kick. forcesAreCurrent = 0 computes F(t) first. md: a handle from ub200_md_create
 * (declared below). */
struct ub200_md;
int ub200_md_lj_nve_verlet_run_f32(struct ub200_md *md, ub200_verletlist *vl, void *d_pos, void *d_vel, void *d_force, int N,
                                   const float L[3], float rc, const float *params, int ntypes, float dt, int nsteps,
                                   int forcesAreCurrent, void *stream);

/* DPD transverser (Potential/DPD.cuh:92-159). d_vel: real3[*] indexed by GLOBAL index like getInfo(pi).
 * sigma = sqrt(2 T)/sqrt(dt) as DPD_impl computes it (:66,:84-92); seed/step are the Saru seeds (:129).
 * idStride = N used in ij = min + N*max (int32 arithmetic, wraps like the reference). */
int ub200_dpd_sum_f32(ub200_celllist *cl, const void *d_vel, float A, float gamma, float sigma, float rcut,
                      uint32_t seed, uint32_t step, int idStride, void *d_force, const int *d_globalIdx,
                      void *stream);

/* Multi-GPU particle decomposition of the DPD forces (BASELINE config 4 shape): every rank holds all positions and
 * velocities, forces only for the particles whose index lies in [ownerLo, ownerHi). The noise is keyed on the global
 * pair indices, so the result does not depend on the number of ranks. accumulate = 0 writes (x,y,z,0), 1 adds. */
int ub200_dpd_sum_owned_f32(ub200_celllist *cl, const void *d_vel, float A, float gamma, float sigma, float rcut,
                            uint32_t seed, uint32_t step, int idStride, void *d_force, int ownerLo, int ownerHi,
                            int accumulate, void *stream);

/* Brick decomposition of the DPD forces: the local arrays hold owned particles [ownerLo, ownerHi) followed by ghosts;
 * d_noiseId[i] is the GLOBAL id of local particle i, used only in the Saru key ij = min + idStride*max (DPD.cuh:128),
 * with idStride = global N, so the pairwise noise is the single-GPU one whatever the decomposition. */
int ub200_dpd_sum_owned_ids_f32(ub200_celllist *cl, const void *d_vel, float A, float gamma, float sigma, float rcut,
                                uint32_t seed, uint32_t step, int idStride, void *d_force, int ownerLo, int ownerHi,
                                int accumulate, const int *d_noiseId, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Path 1e: Langevin velocity Verlet (Gronbech-Jensen & Farago). Replaces VerletNVT::GronbechJensen_ns::integrateGPU<1|2>
 * (Integrator/VerletNVT/GronbechJensen.cu:30-66), the integrator generic_md and examples/misc/benchmark.cu drive.
 * step 1 moves positions and velocities with the noise Saru(index in group, stepNum, seed) and ZEROES the forces, step 2
 * is the closing half kick. noiseAmplitude = sqrt(2 dt friction T) as VerletNVT::Basic computes it (Basic.cu:45);
 * defaultMass > 0 overrides d_mass (Basic.cu:46-50). Bit-identical to the reference kernel.
 * ub200_nvt_initial_velocities_f32 replaces Basic_ns::initialVelocities (Basic.cu:12-29, velAmplitude = sqrt(3 T)).
 * ------------------------------------------------------------------------------------------------ */
int ub200_nvt_gj_half_step_f32(void *d_pos, void *d_vel, void *d_force, const float *d_mass, float defaultMass,
                               const int *d_groupIdx, int N, float dt, float friction, int is2D, float noiseAmplitude,
                               uint32_t stepNum, uint32_t seed, int step, void *stream);
/* VerletNVT::Basic_ns::integrateGPU<1|2> (Integrator/VerletNVT/Basic.cu:87-117): the plain Langevin velocity Verlet, same
 * arguments as ub200_nvt_gj_half_step_f32; BOTH half steps kick with friction and a fresh noise draw
 * (Saru(index + N (step - 1), stepNum, seed)), step 1 also drifts and zeroes the force. Bit-identical to the reference. */
int ub200_nvt_basic_half_step_f32(void *d_pos, void *d_vel, void *d_force, const float *d_mass, float defaultMass,
                                  const int *d_groupIdx, int N, float dt, float friction, int is2D, float noiseAmplitude,
                                  uint32_t stepNum, uint32_t seed, int step, void *stream);
int ub200_nvt_initial_velocities_f32(void *d_vel, const int *d_groupIdx, int N, float velAmplitude, int is2D, uint32_t seed,
                                     void *stream);

/* ------------------------------------------------------------------------------------------------
 * Path 1c: velocity Verlet. Replaces VerletNVE_ns::integrateGPU<1|2> (Integrator/VerletNVE.cu:64-85)
 * and VerletNVE::resetForces (:152-158). d_mass may be NULL (defaultMass used). step==1 also drifts.
 * ------------------------------------------------------------------------------------------------ */
int ub200_nve_half_step_f32(void *d_pos, void *d_vel, const void *d_force, const float *d_mass,
                            float defaultMass, const int *d_groupIdx, int N, float dt, int is2D, int step,
                            void *stream);

/* second kick of step n fused with kick + drift of step n+1 (unit mass, all particles [0,N)): same roundings as
 * ub200_nve_half_step_f32(step=2) followed by (step=1) */
int ub200_nve_kick_kick_drift_f32(void *d_pos, void *d_vel, const void *d_force, int N, float dt, void *stream);

/* Fused engine for a whole VerletNVE::forwardTime with one PairForces<LJ,CellList> interactor
 * (Integrator/VerletNVE.cu:174-188): kick+drift, cell list rebuild, LJ forces (written, not accumulated),
 * second kick. d_force must hold F(t) on entry (see ub200_md_lj_nve_prepare_f32). nsteps steps are
 * enqueued back to back on the stream. */
typedef struct ub200_md ub200_md;
int ub200_md_create(ub200_md **out);
int ub200_md_destroy(ub200_md *md);
int ub200_md_lj_nve_prepare_f32(ub200_md *md, void *d_pos, void *d_force, int N, const float L[3], float rc,
                                const float *params, int ntypes, void *stream);
int ub200_md_lj_nve_run_f32(ub200_md *md, void *d_pos, void *d_vel, void *d_force, int N, const float L[3],
                            float rc, const float *params, int ntypes, float dt, int nsteps, void *stream);
/* same through HOST buffers (pinned or pageable): H2D pos+vel, prepare, nsteps, D2H pos+vel+force; synchronises. */
int ub200_md_lj_nve_run_host_f32(ub200_md *md, float *h_pos4, float *h_vel3, float *h_force4, int N,
                                 const float L[3], float rc, const float *params, int ntypes, float dt,
                                 int nsteps, void *stream);
/* Observable of a run: kinetic energy sum v^2 / 2 (unit mass, double accumulation, fixed summation order) of d_vel
 * real3[N], copied asynchronously into *h_out (pinned host memory for a truly asynchronous copy); the value is valid
 * once the stream has been synchronised. */
int ub200_md_kinetic_energy_f32(ub200_md *md, const void *d_vel, int N, double *h_out, void *stream);
struct ub200_ljengine *ub200_md_engine(ub200_md *md); /* the pair-force engine the fused loop drives (declared below) */
/* ------------------------------------------------------------------------------------------------
 * Path 2: FFT-based hydrodynamics. Precision is chosen at create time (precisionBytes = 4 | 8, the
 * reference's global `real`); positions/forces are real4, per-particle outputs real3, grids real3 AoS
 * [nz][ny][nxPad][3] with nxPad = 2(nx/2+1) which the FFT turns IN PLACE into complex3 AoS
 * [nz][ny][nx/2+1][3] - the reference's cufftMakePlanMany(batch 3, stride 3) layout
 * (Integrator/BDHI/FCM/FCM_impl.cuh:186-211).
 * ------------------------------------------------------------------------------------------------ */
#define UB200_KERNEL_PESKIN3 0  /* IBM_kernels::Peskin::threePoint  misc/IBM_kernels.cuh:118-137 */
#define UB200_KERNEL_PESKIN4 1  /* IBM_kernels::Peskin::fourPoint   misc/IBM_kernels.cuh:140-157 */
#define UB200_KERNEL_GAUSSIAN 2 /* FCM_ns::Kernels::Gaussian        Integrator/BDHI/FCM/FCM_kernels.cuh:22-58 */
#define UB200_KERNEL_BARNETT_MAGLAND 3 /* IBM_kernels::BarnettMagland          misc/IBM_kernels.cuh:83-113 */
#define UB200_KERNEL_SIXPOINT 4        /* IBM_kernels::GaussianFlexible::sixPoint misc/IBM_kernels.cuh:163-237 */
typedef struct {
  int kind;
  int support;  /* points per dimension (Kernel::support / getMaxSupport) */
  double h;     /* Peskin, six point: grid spacing */
  double prefactor, tau, rmax; /* Gaussian: prefactor*exp(tau r^2) for r < rmax.
                                  Barnett-Magland: prefactor = 1/norm (= phi(0)), tau = beta, rmax = alpha (half support):
                                  phi(r) = exp(beta (sqrt(1 - (r/alpha)^2) - 1)) / norm for |r| <= alpha */
} ub200_ibm_kernel;

/* 3-D real FFT, hand written (no cuFFT). Replaces the cuFFT plans + cufftExecR2C/D2Z/C2R/Z2D calls
 * (FCM_impl.cuh:179-234,293-304,544-557; PSE/FarField.cuh:555-603). Unnormalised; direction -1 = forward
 * (real -> complex), +1 = inverse. Sizes with prime factors 2, 3, 5, 7, 11. */
typedef struct ub200_fft3d ub200_fft3d;
int ub200_fft3d_create(ub200_fft3d **out, int precisionBytes, int nx, int ny, int nz);
int ub200_fft3d_destroy(ub200_fft3d *plan);
int ub200_fft3d_exec(ub200_fft3d *plan, void *d_grid, int direction, void *stream);

/* Immersed boundary spreading / interpolation. Replaces IBM<Kernel>::spread / gather (misc/IBM.cuh:117-184,
 * kernels misc/IBM.cu:83-147,168-235). spread ADDS into d_grid3 and gather ADDS into d_out3 like the
 * reference; spread_overwrite writes every node of the grid (no prior zero fill needed) and is what the
 * FCM/PSE pipelines use. d_val: real3 (valStride 3) or real4 (valStride 4) per particle. */
typedef struct ub200_ibm ub200_ibm;
int ub200_ibm_create(ub200_ibm **out, int precisionBytes, const double L[3], const int periodic[3], const int cells[3],
                     const ub200_ibm_kernel *kernel, int nxPad);
int ub200_ibm_destroy(ub200_ibm *ibm);
int ub200_ibm_spread(ub200_ibm *ibm, const void *d_pos, const void *d_val, int valStride, int N, void *d_grid3,
                     void *stream);
int ub200_ibm_spread_overwrite(ub200_ibm *ibm, const void *d_pos, const void *d_val, int valStride, int N,
                               void *d_grid3, void *stream);
int ub200_ibm_gather(ub200_ibm *ibm, const void *d_pos, int N, const void *d_grid3, void *d_out3, void *stream);

/* Force Coupling Method. Replaces FCM_impl<Kernel,KernelTorque>::computeHydrodynamicDisplacements without
 * torques (Integrator/BDHI/FCM/FCM_impl.cuh:652-693) and hence BDHI::FCM::computeMF (BDHI_FCM.cuh:131-142):
 * d_out3[i] = (M F)_i + prefactor*sqrt(2 T)*(M^1/2 dW)_i. d_force may be NULL (noise only). The Brownian
 * noise follows fourierBrownianNoise (FCM_impl.cuh:437-542): Saru(node id, seed, call counter). */
typedef struct ub200_fcm ub200_fcm;
int ub200_fcm_create(ub200_fcm **out, int precisionBytes, const double L[3], const int cells[3],
                     const ub200_ibm_kernel *kernel, double viscosity, uint32_t seed);
int ub200_fcm_destroy(ub200_fcm *fcm);
int ub200_fcm_mdot(ub200_fcm *fcm, const void *d_pos, const void *d_force, int N, double temperature,
                   double prefactor, void *d_out3, void *stream);
/* Rotational FCM. Replaces the torque path of FCM_impl::computeHydrodynamicDisplacements (FCM_impl.cuh:306-358,583-649,
 * 652-693): torques (real4) are spread with KernelTorque (FCM_ns::Kernels::GaussianTorque, FCM_kernels.cuh:60-80; set once
 * with ub200_fcm_set_torque_kernel), 1/2 i dk x T is added to the force spectrum before the Stokes operator, and the
 * angular velocities 1/2 curl v are interpolated with KernelTorque. d_force may be NULL. */
int ub200_fcm_set_torque_kernel(ub200_fcm *fcm, const ub200_ibm_kernel *kernelTorque);
int ub200_fcm_mdot_torque(ub200_fcm *fcm, const void *d_pos, const void *d_force, const void *d_torque, int N,
                          double temperature, double prefactor, void *d_linear3, void *d_angular3, void *stream);
int ub200_fcm_grid_info(ub200_fcm *fcm, int cells[3], int *nxPad, void **d_grid);
/* BDHI::EulerMaruyama position update. Replaces EulerMaruyama_ns::integrateGPUD
 * (Integrator/BDHI/BDHI_EulerMaruyama.cu:82-113): x += dt (K x + MF) + sqrt2Tdt BdW. d_MF/d_BdW real3[N] indexed by
 * group slot, d_BdW and K9 (host, row major 3x3 shear matrix) may be NULL. */
int ub200_bdhi_euler_update(int precisionBytes, void *d_pos, const int *d_groupIdx, const void *d_MF, const void *d_BdW,
                            const double *K9, int N, double sqrt2Tdt, double dt, int is2D, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU FCM (new functionality: the reference is single-GPU). One process per GPU of one
 * NVSwitch box; the grid is decomposed in z slabs, the 3-D FFT transposes through peer-mapped NVLink stores
 * fused into the y / z passes (no NCCL on the data path), halo planes of the interpolation are read through
 * peer-mapped pointers. Same semantics as ub200_fcm_mdot, every rank passes the same (replicated) positions
 * and forces and receives the full result; the result is bit-identical to the single-GPU one.
 * Set-up: create on every rank, exchange the ub200_fcm_dist_ipc_size()-byte blobs of ub200_fcm_dist_ipc_export
 * between the ranks (any transport; rank order), hand all of them to ub200_fcm_dist_ipc_import.
 * cells[1] and cells[2] must be divisible by world (<= 8); Peskin 3/4-point kernels.
 * ------------------------------------------------------------------------------------------------ */
typedef struct ub200_fcm_dist ub200_fcm_dist;
int ub200_fcm_dist_create(ub200_fcm_dist **out, int precisionBytes, const double L[3], const int cells[3],
                          const ub200_ibm_kernel *kernel, double viscosity, uint32_t seed, int rank, int world,
                          int maxParticles);
int ub200_fcm_dist_destroy(ub200_fcm_dist *fcm);
int ub200_fcm_dist_ipc_size(void);
int ub200_fcm_dist_ipc_export(ub200_fcm_dist *fcm, void *blob);
int ub200_fcm_dist_ipc_import(ub200_fcm_dist *fcm, const void *blobsOfAllRanks);
int ub200_fcm_dist_mdot(ub200_fcm_dist *fcm, const void *d_pos, const void *d_force, int N, double temperature,
                        double prefactor, void *d_out3, void *stream);
/* PSE far field on the same slab machinery (BASELINE config 3: "slab-decomposed FFT over 8 GPUs"): switch the spectral
 * operator of a handle created with the PSE Gaussian window (supports 3, 5, 7) to the Hasimoto-split RPY Green's function
 * (FarField.cuh:85-153). seedFar / seed2 as in ub200_pse_far_mdot; noise prefactor = prefactor sqrt(2 T / dV) (:467-492).
 * ub200_fcm_dist_mdot then returns Mw F + noise (overwriting d_out3; the caller accumulates like IBM::gather would). */
int ub200_fcm_dist_set_pse_operator(ub200_fcm_dist *fcm, double hydrodynamicRadius, double psi, double eta, double shearStrain);
int ub200_fcm_dist_set_noise_seed2(ub200_fcm_dist *fcm, uint32_t seed2);
/* diagnostics (environment UB200_DIST_PROFILE=1, makes every call synchronous): mean ms of the 12 phases of mdot */
int ub200_fcm_dist_profile(ub200_fcm_dist *fcm, double phases[12]);
/* synchronises the stream; *flag != 0 when a peer barrier timed out (a rank died) */
int ub200_fcm_dist_error_flag(ub200_fcm_dist *fcm, void *stream, int *flag);

/* ------------------------------------------------------------------------------------------------
 * Positively Split Ewald RPY hydrodynamics. Replaces BDHI::PSE (Integrator/BDHI/BDHI_PSE.cuh:82-176) =
 * pse_ns::FarField (PSE/FarField.cuh:317-553) + pse_ns::NearField (PSE/NearField.cuh:29-282) + the Lanczos
 * square root (misc/LanczosAlgorithm/LanczosAlgorithm.cu:27-250). ub200_pse_create resolves every derived
 * parameter exactly like the reference constructors: near-field cut-off and F/G table (NearField.cuh:65-102),
 * far-field grid through nextFFTWiseSize3D (FarField.cuh:646-654, utils/Grid.cuh:142-213), Gaussian support
 * and eta (FarField.cuh:605-644). seedNear / seedFar are the two sys->rng().next32() draws the reference
 * makes at construction (near first: initialization.cu:57-59); seed2 arguments are the per-call draws
 * (FarField.cuh:478, NearField.cuh:274).
 * ------------------------------------------------------------------------------------------------ */
typedef struct ub200_pse ub200_pse;
typedef struct {
  double L[3];
  double viscosity, hydrodynamicRadius, tolerance, psi, shearStrain; /* pse_ns::Parameters (PSE/utils.cuh:17-24) */
  int cellsOverride[3]; /* {0,0,0}: grid chosen like the reference; otherwise forced (tests on small grids) */
} ub200_pse_params;
typedef struct {
  int cells[3], support, nTable, lastLanczosIterations;
  double eta, rcut;
  const void *d_table; /* real2[nTable]: F, G divided by 6 pi eta a */
  void *d_grid;
  ub200_ibm_kernel kernel; /* the far-field Gaussian window (pse_ns::Kernel, FarField.cuh:25-41) as resolved by create */
} ub200_pse_info_t;
int ub200_pse_create(ub200_pse **out, int precisionBytes, const ub200_pse_params *par, uint32_t seedNear, uint32_t seedFar);
int ub200_pse_destroy(ub200_pse *pse);
int ub200_pse_info(ub200_pse *pse, ub200_pse_info_t *info);
int ub200_pse_set_shear_strain(ub200_pse *pse, double strain); /* PSE::setShearStrain BDHI_PSE.cuh:160-165 */
/* FarField::computeHydrodynamicDisplacements (FarField.cuh:535-553): d_MF3 += Mw F + prefactor sqrt(2T) Mw^1/2 dW.
 * d_force (real4) may be NULL (noise only). */
int ub200_pse_far_mdot(ub200_pse *pse, const void *d_pos, const void *d_force, int N, double temperature, double prefactor,
                       uint32_t seed2, void *d_MF3, void *stream);
/* NearField::Mdot (NearField.cuh:243-252): rebuilds the cell list, d_Mv3 += Mr v. d_v: real4 (vStride 4) or real3 (3). */
int ub200_pse_near_mdot(ub200_pse *pse, const void *d_pos, const void *d_v, int vStride, int N, void *d_Mv3, void *stream);
/* NearField::computeStochasticDisplacements (NearField.cuh:254-282): d_BdW3 = prefactor sqrt(2T) Mr^1/2 dW (overwritten,
 * like the reference's gemv with beta = 0). Host-synchronous (Lanczos convergence checks) like the reference. */
/* The same product through the near field's Verlet list, for a step that also draws noise: the Lanczos iteration needs the list
 * anyway (it walks it 5 - 10 times), so it is built here and ub200_pse_near_noise_reuse skips its own build when it is given the
 * SAME position array, unchanged since (the reference's BDHI::EulerMaruyama calls computeMF and computeBdW back to back). */
int ub200_pse_near_mdot_list(ub200_pse *pse, const void *d_pos, const void *d_v, int vStride, int N, void *d_Mv3, void *stream);
int ub200_pse_near_noise_reuse(ub200_pse *pse, const void *d_pos, int N, double temperature, double prefactor, uint32_t seed2,
                               void *d_BdW3, int *iterations, void *stream);
int ub200_pse_near_noise(ub200_pse *pse, const void *d_pos, int N, double temperature, double prefactor, uint32_t seed2,
                         void *d_BdW3, int *iterations, void *stream);
/* same, ADDED to d_out3: what PSE::computeHydrodynamicDisplacements (BDHI_PSE.cuh:141-158) documents - "Mobility force +
 * prefactor sqrt(2 T M) dW" - needs. The reference passes MF itself to the Lanczos solver, whose final gemv has beta = 0,
 * so its near-field M F is overwritten whenever T > 0 and a force is given; this entry point keeps both terms. */
int ub200_pse_near_noise_add(ub200_pse *pse, const void *d_pos, int N, double temperature, double prefactor, uint32_t seed2,
                             void *d_out3, int *iterations, void *stream);

/* ---- PSE near field and its Lanczos noise over the ranks of one NVSwitch domain (the reference is single
 * GPU: NearField::Mdot NearField.cuh:236-251, computeStochasticDisplacements :254-282, lanczos::Solver::run
 * LanczosAlgorithm.cu:202-228). Positions (and the force vector of Mdot) are replicated; rank r owns the rows
 * [r N / world, (r + 1) N / world) of the cell-sorted order. A product evaluates only the owned rows; the next Krylov
 * vector is stored straight into the records of every rank (peer stores) and the two scalars of an iteration are summed
 * over the ranks by a one-warp kernel that is also the barrier - in rank order, so every rank sees the same bits and
 * takes the same convergence decisions. Results land, complete, on every rank (added to d_Mv3 / d_out3).
 * Set-up like ub200_brick_*: ub200_pse_dist_create on every rank, exchange the ub200_comm_ipc_size()-byte blobs of
 * ub200_pse_dist_ipc_export and hand all of them (rank order) to ub200_pse_dist_ipc_import; ranks that share a process
 * (tests) exchange ub200_pse_dist_arena pointers through ub200_pse_dist_attach_local instead. All ranks must then make
 * the same sequence of near_mdot / near_noise_add calls. */
int ub200_pse_dist_create(ub200_pse *pse, int rank, int world, int maxParticles);
int ub200_pse_dist_ipc_export(ub200_pse *pse, void *blob);
int ub200_pse_dist_ipc_import(ub200_pse *pse, const void *blobsOfAllRanks);
int ub200_pse_dist_arena(ub200_pse *pse, void **arena);
int ub200_pse_dist_attach_local(ub200_pse *pse, void *const *arenasOfAllRanks);
/* neighbour list of the replicated positions (every rank builds it; no communication) */
int ub200_pse_dist_near_prepare(ub200_pse *pse, const void *d_pos, int N, void *stream);
/* d_Mv3 += M_near v (v replicated: real3 or real4 rows, vStride 3 or 4) */
int ub200_pse_dist_near_mdot(ub200_pse *pse, const void *d_v, int vStride, int N, void *d_Mv3, void *stream);
/* d_out3 += prefactor sqrt(2 T) M_near^1/2 dW; the noise of a particle is keyed by its index as in ub200_pse_near_noise,
 * so the result equals the single-GPU one up to the summation order of the dot products */
int ub200_pse_dist_near_noise_add(ub200_pse *pse, int N, double temperature, double prefactor, uint32_t seed2, void *d_out3,
                                  int *iterations, void *stream);
/* reads back (synchronising the stream) whether a peer barrier ever timed out */
int ub200_pse_dist_error_flag(ub200_pse *pse, void *stream, int *flag);

/* ------------------------------------------------------------------------------------------------
 * BASELINE config 0: BD::EulerMaruyama (ideal or with interactor forces). Replaces EulerMaruyama_ns::integrateGPU
 * (Integrator/BrownianDynamics.cu:117-145, launched by EulerMaruyama::updatePositions :158-173):
 *   M = selfMobility * (radius ? 1/radius[i] : 1);  R += dt (K R + M F);  R += (gf(0,B).x, gf(0,B).y, gf'(0,B).x),
 *   B = sqrt(2 T M dt), gf from Saru(i, stepNum, seed) with i the particle (global) index.
 * d_force: real4[*] or NULL (ideal particles); K9: host row-major shear matrix or NULL; d_radius: real[*] or NULL.
 * The caller increments stepNum before every call like EulerMaruyama::forwardTime does (:148-155); seed is the
 * integrator's Saru seed (BaseBrownianIntegrator ctor :14-16). Results are bit-identical to the reference's.
 * ------------------------------------------------------------------------------------------------ */
int ub200_bd_euler_maruyama_step(int precisionBytes, void *d_pos, const int *d_groupIdx, const void *d_force,
                                 const double *K9, double selfMobility, const void *d_radius, double dt, int is2D,
                                 double temperature, int N, uint32_t stepNum, uint32_t seed, void *stream);

/* number of kernel launches the library enqueued since process start (bench.py's gpu_launches) */
unsigned long long ub200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
