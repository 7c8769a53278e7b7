// C-ABI surface that is not tied to one kernel family (include/egotap_b200.h).
#include "host_util.cuh"

extern "C" int egotap_b200_abi_version(void) { return EGOTAP_B200_ABI_VERSION; }
extern "C" const char* egotap_b200_last_error(void) { return eb::err_buf(); }
extern "C" long long egotap_b200_launch_count(void) { return eb::launch_counter().load(); }

#include <vector>
namespace eb {
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
bool& prof_on() { return g_prof_on; }
void prof_push(const ProfRec& r) { g_prof.push_back(r); }
}  // namespace eb

extern "C" int egotap_b200_profile_begin(void) {
  for (auto& r : eb::g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  eb::g_prof.clear();
  eb::g_prof_on = true;
  return 0;
}
extern "C" int egotap_b200_profile_end(int* num_records) {
  eb::g_prof_on = false;
  EB_CUDA(cudaDeviceSynchronize());
  if (num_records) *num_records = int(eb::g_prof.size());
  return 0;
}
extern "C" int egotap_b200_profile_record(int i, const char** name, int* M, int* N, int* K, int* groups, int* variant,
                                          float* ms) {
  if (i < 0 || i >= int(eb::g_prof.size())) return eb::fail(EGOTAP_E_ARG, "profile_record: index %d out of range", i);
  const eb::ProfRec& r = eb::g_prof[i];
  *name = r.name; *M = r.M; *N = r.N; *K = r.K; *groups = r.groups; *variant = r.variant;
  EB_CUDA(cudaEventElapsedTime(ms, r.e0, r.e1));
  return 0;
}
