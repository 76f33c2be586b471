// NG*NG <= 512 sort width with the number of g-ordinates fixed to 20 at compile time
#include "koverlap_impl.cuh"
int ov_dispatch_16_ng20(const OvParams &P, bool grad, cudaStream_t stream) { return ov_dispatch_np<16, 20>(P, grad, stream); }
