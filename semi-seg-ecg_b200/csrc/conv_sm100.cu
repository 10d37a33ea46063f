// placeholder, replaced below
#include "common.cuh"
int ssb_conv1d_fwd_sm100(const void*, const void*, void*, ssb_geom, ssb_geom, int, int, cudaStream_t) { ssb_set_error("tcgen05 conv not built"); return SSB_ERR_UNSUPPORTED; }
int ssb_conv1d_dgrad_sm100(const void*, const void*, void*, ssb_geom, ssb_geom, int, int, int, cudaStream_t) { ssb_set_error("tcgen05 conv not built"); return SSB_ERR_UNSUPPORTED; }
int ssb_conv1d_wgrad_sm100(const void*, const void*, float*, ssb_geom, ssb_geom, int, int, cudaStream_t) { ssb_set_error("tcgen05 conv not built"); return SSB_ERR_UNSUPPORTED; }
