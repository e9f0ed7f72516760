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
    static const KTable<float> t = {&launch_kprobe<float, BF_NB>, &launch_magfit<float, BF_NB>, &launch_refit<float, BF_NB>,
                                    &launch_flux<float, BF_NB>,
                                    &launch_records<float, BF_NB, float>,
                                    &launch_records<float, BF_NB, double>};
    return &t;
}
const KTable<double>* BF_CAT(ktable_f64_, BF_NB)() {
    static const KTable<double> t = {&launch_kprobe<double, BF_NB>, &launch_magfit<double, BF_NB>, &launch_refit<double, BF_NB>,
                                     &launch_flux<double, BF_NB>,
                                     &launch_records<double, BF_NB, double>,
                                     &launch_records<double, BF_NB, double>};
    return &t;
}
}  // namespace bf
