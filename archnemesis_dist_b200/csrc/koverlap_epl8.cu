// koverlap_epl8.cu -- instantiates the overlap kernels that keep 8 sort keys per lane (NG*NG <= 256).
#include "koverlap_impl.cuh"
int ov_dispatch_8(const OvParams &P, bool grad, cudaStream_t stream) { return ov_dispatch_np<8>(P, grad, stream); }
