// C-ABI surface that is not tied to one kernel family (include/egotap_b200.h).
#include "host_util.cuh"

extern "C" int egotap_b200_abi_version(void) { return EGOTAP_B200_ABI_VERSION; }
extern "C" const char* egotap_b200_last_error(void) { return eb::err_buf(); }
extern "C" long long egotap_b200_launch_count(void) { return eb::launch_counter().load(); }
