// K3 (tensor-core path): placeholder until the tcgen05 digit-plane kernel lands.
#include "snn.cuh"

namespace lens {

bool snn_tc_supported(const SnnHandle *) { return false; }
int snn_tc_prepare(SnnHandle *, cudaStream_t) { set_err("tensor-core path not built"); return -1; }
void snn_tc_release(SnnHandle *) {}
int snn_tc_output(SnnHandle *, const int8_t *, int, int, int, float *, uint8_t *, cudaStream_t)
{
    set_err("tensor-core path not built");
    return -1;
}

}  // namespace lens
