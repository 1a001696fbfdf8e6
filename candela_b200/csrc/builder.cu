#include "builder.cuh"

namespace cndl {

int build_object(BuildRequest& rq, cudaStream_t st, LaunchCounter& lc, float* build_ms, std::string& err) {
    (void)rq; (void)st; (void)lc; (void)build_ms;
    err = "GPU builder not implemented yet";
    return CNDL_ERR_INVALID;
}

}  // namespace cndl
