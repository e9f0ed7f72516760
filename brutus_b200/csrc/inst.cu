// inst.cu -- compiled once per band count (-DBF_NB=n): instantiates the band-templated kernels for
// float32 (throughput) and float64 (verification) and exports their launch table.
#include "kernels_nb.cuh"

#ifndef BF_NB
#error "compile with -DBF_NB=<number of bands>"
#endif
#define BF_CAT2(a, b) a##b
#define BF_CAT(a, b) BF_CAT2(a, b)

namespace bf {
const KTable<float>* BF_CAT(ktable_f32_, BF_NB)() {
    static const KTable<float> t = {&launch_kprobe<float, BF_NB>, &launch_sweep<float, BF_NB>, &launch_fixup<float, BF_NB>,
                                    &launch_flux_more<float, BF_NB>};
    return &t;
}
const KTable<double>* BF_CAT(ktable_f64_, BF_NB)() {
    static const KTable<double> t = {&launch_kprobe<double, BF_NB>, &launch_sweep<double, BF_NB>, &launch_fixup<double, BF_NB>,
                                     &launch_flux_more<double, BF_NB>};
    return &t;
}
}  // namespace bf
