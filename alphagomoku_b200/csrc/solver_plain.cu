// K5 for the rule sets without forbidden moves (freestyle, standard, caro): the same sources as solver.cu's kernel with every renju branch
// compiled out (solver_kernel.cuh).
#define AGB_SOLVER_NO_RENJU
#define AGB_SOLVER_NS solver_plain
#include "solver_kernel.cuh"
