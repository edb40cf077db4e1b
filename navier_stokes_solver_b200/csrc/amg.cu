// amg.cu -- smoothed-aggregation AMG V-cycle on the velocity block (stand-in for Trilinos ML through
// TrilinosWrappers::PreconditionAMG, NSSolverStationary.hpp:225).
#include "device.cuh"

namespace nsx {

struct AmgHierarchy {};

void amg_setup(Ctx &, const DevCSR &) { throw std::logic_error("AMG is not built yet"); }
void amg_apply(Ctx &, double *, const double *) { throw std::logic_error("AMG is not built yet"); }

}  // namespace nsx
