// Handle of the LJ pair-force engine: PairForces<Potential::LJ, CellList>::sum (Interactor/PairForces.cu:43-78) in one call.
#pragma once
#include "common.cuh"
#include "../../include/uammd_b200/colgeom.h"

// Engine-private half-cell list (see colgeom.h): particles sorted by linear half-cell index (x fastest), stable inside a
// cell, coordinates folded into the primary box and made consistent with the cell (a coordinate that rounds into cell n
// is stored one box length lower, in cell 0).
struct ub200_ljengine {
  ub200::GridF grid;            // the half-cell grid in the reference's fp32 arithmetic
  ub200::ColGrid cg;
  int N = 0, ncells = 0, binCells = 0;
  ub200::DevBuf pos, idx;       // float4[N] canonical positions (w = type), int[N] sorted slot -> group index
  ub200::DevBuf binCount, binStart, blockSums, codeSlot, unstable, errorFlag;
  ub200_celllist *cl = nullptr; // reference-layout list for the grids the column traversal does not take
  ub200::LJTableCache table;
  int lastPath = -1;            // 0 column traversal, 1 cell traversal, 2 all pairs
};

namespace ub200 {
// forces (and optionally energies / virials) of all particles; see ub200_ljengine_sum_f32
int ljEngineSum(ub200_ljengine *e, const float4 *pos, const int *groupIdx, int N, const float L[3], const int periodic[3],
                const float *params, int ntypes, float4 *force, float *energy, float *virial, const int *globalIdx,
                bool accumulate, int ownerLo, int ownerHi, cudaStream_t st);
// multi-GPU bricks (brick_md.cu): list of a rank's local arrays [owned | ghosts] (*nLocalDev particles, at most maxN) on its
// window cg of the global half-cell grid, cells ordered by sortKey (global ids); then the forces of the owned block
int ljEngineBuildWindow(ub200_ljengine *e, const float4 *pos, const int *sortKey, int maxN, const int *nLocalDev,
                        const float L[3], const int periodic[3], const int globalDims[3], const ColGrid &cg, cudaStream_t st);
int ljEngineTraverseWindow(ub200_ljengine *e, const int *nOwnedDev, const float *params, int ntypes, float4 *force,
                           bool accumulate, cudaStream_t st);
// half cells per dimension for this box and cut-off; false when the column traversal does not apply
bool ljEngineDims(const float L[3], const int periodic[3], float rc, int dims[3], int per[3]);
} // namespace ub200
