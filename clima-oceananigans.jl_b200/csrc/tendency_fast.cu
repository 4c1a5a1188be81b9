// tendency_fast.cu -- specialised tendency kernels for the headline configuration.
#include "internal.h"

namespace ob {

template <class FT>
bool launch_tendency_fast(const Phys<FT>& P, int comp, const FT* const U[3], const FT* psi,
                          const FT* pHY, FT* Gn, const FT* Gm, FT* psi_new, const Substep<FT>& ss) {
    return false;
}
template bool launch_tendency_fast<float>(const Phys<float>&, int, const float* const[3], const float*,
                                          const float*, float*, const float*, float*, const Substep<float>&);
template bool launch_tendency_fast<double>(const Phys<double>&, int, const double* const[3], const double*,
                                           const double*, double*, const double*, double*, const Substep<double>&);
}  // namespace ob
