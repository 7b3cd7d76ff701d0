// TEST INFRASTRUCTURE ONLY -- CPU oracle (see mo_array.h).
// Kernels whose restatement is not written yet fail loudly.
#include "mo_kernels.h"
namespace mo {
#ifndef MO_HAVE_BDS
void bds_box(const mgpu_params&, const Arr&, Arr*, const Arr*, const Arr&, const int*, const int*, int, bool) {
  fail("oracle: bds not restated yet");
}
#endif
#ifndef MO_HAVE_VELPRED
void mkutrans_box(const mgpu_params&, const Arr&, const Arr&, Arr*, const double*, const int*, const int*,
                  const int*, const int*, int, const Arr*) {
  fail("oracle: mkutrans not restated yet");
}
void velpred_box(const mgpu_params&, const Arr&, const Arr&, Arr*, const Arr*, const Arr&, const double*,
                 const int*, const int*, const int*, const int*, int, const Arr*) {
  fail("oracle: velpred not restated yet");
}
#endif
}  // namespace mo
