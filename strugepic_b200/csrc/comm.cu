// placeholder: multi-GPU layer not yet implemented
#include "engine.cuh"
namespace spic {
int comm_exchange_fill(Ctx* c, double*) { c->err = "multi-GPU not built yet"; return SPIC_ENCCL; }
int comm_exchange_sum(Ctx* c, double*, int) { c->err = "multi-GPU not built yet"; return SPIC_ENCCL; }
int comm_collect_leavers(Ctx* c, Species&, double* const*, double* const*, const int*, const unsigned*, unsigned) {
  c->err = "multi-GPU not built yet";
  return SPIC_ENCCL;
}
int comm_migrate(Ctx* c) { c->err = "multi-GPU not built yet"; return SPIC_ENCCL; }
int comm_allreduce_sum(Ctx* c, double*, int) { c->err = "multi-GPU not built yet"; return SPIC_ENCCL; }
void comm_destroy(Ctx*) {}
}  // namespace spic
extern "C" {
int spic_comm_unique_id(void*) { return SPIC_ENCCL; }
int spic_comm_init(spic_ctx*, const void*) { return SPIC_ENCCL; }
}
