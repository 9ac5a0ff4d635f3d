// fp32-storage instantiations of the predict / weighting kernels (see the end of pfb_predict_loglik.cu): a translation
// unit of its own so that the two halves of the largest source compile in parallel.
#define PFB_PL_F32_TU 1
#include "pfb_predict_loglik.cu"
