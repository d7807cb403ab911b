// temporary stubs while kernel families are brought up one by one
#include "ncl_internal.h"
int try_elementwise(const Nest&) { return NCL_ERR_UNSUPPORTED; }
int try_reduce(const Nest&) { return NCL_ERR_UNSUPPORTED; }
int try_transpose(const Nest&) { return NCL_ERR_UNSUPPORTED; }
int try_gemm(const Nest&) { return NCL_ERR_UNSUPPORTED; }
extern "C" {
int ncl_comm_unique_id(void*) { return set_error(NCL_ERR_UNSUPPORTED, "comm not built"); }
int ncl_comm_init(const void*, int, int) { return set_error(NCL_ERR_UNSUPPORTED, "comm not built"); }
int ncl_comm_destroy(void) { return NCL_OK; }
int ncl_allreduce(int, ncl_tensor*) { return set_error(NCL_ERR_UNSUPPORTED, "comm not built"); }
int ncl_reduce_all_sharded(int, const ncl_tensor*, ncl_tensor*) { return set_error(NCL_ERR_UNSUPPORTED, "comm not built"); }
}
