// koverlap_epl16.cu -- instantiates the overlap kernels that keep 16 sort keys per lane (NG*NG <= 512).
#include "koverlap_impl.cuh"
int ov_dispatch_16(const OvParams &P, bool grad, cudaStream_t stream) { return ov_dispatch_np<16>(P, grad, stream); }
