// patch_embed.cu -- image patch embedding written into the early-fusion concat buffer (sm_100a).
#include "p3p_internal.cuh"

namespace p3p {

int launch_patch_embed(const float*, int, int, int, int, int, const float*, const float*, int, int, void*, int, int, int,
                       cudaStream_t) {
    return fail(P3P_ERR_UNSUPPORTED, "patch embed kernel not built yet");
}

}  // namespace p3p
