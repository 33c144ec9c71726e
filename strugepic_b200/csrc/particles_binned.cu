// placeholder: binned engine not yet implemented -- everything stays in the direct list
#include "engine.cuh"
namespace spic {
int engine_ingest(Ctx*, Species&) { return SPIC_OK; }
void engine_free_species(Ctx*, Species&) {}
void engine_destroy(Ctx*) {}
int engine_count(Ctx*, Species&, long* nb) { *nb = 0; return SPIC_OK; }
int engine_gather(Ctx*, Species&, double**, double**, long* nb) { *nb = 0; return SPIC_OK; }
int engine_theta_axis(Ctx*, Species&, int, double) { return SPIC_OK; }
int engine_push_v_e(Ctx*, Species&, double) { return SPIC_OK; }
int engine_kinetic(Ctx*, Species&, double*) { return SPIC_OK; }
int engine_deposit_rho(Ctx*, Species&, double*) { return SPIC_OK; }
int engine_set_option(Ctx* c, const char*, double) { c->err = "unknown option"; return SPIC_EINVAL; }
}  // namespace spic
