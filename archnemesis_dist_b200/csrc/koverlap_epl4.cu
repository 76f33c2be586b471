// koverlap_epl4.cu -- instantiates the overlap kernels that keep 4 sort keys per lane (NG*NG <= 128).
#include "koverlap_impl.cuh"
int ov_dispatch_4(const OvParams &P, bool grad, cudaStream_t stream) { return ov_dispatch_np<4>(P, grad, stream); }
