// C entry points over the host mirror's pure helpers (no engine needed), for CPU-side tests.
#include "../../gateway-api-inference-extension_b200/host/epp_scheduler.hpp"

extern "C" int epp_count_fields(const char* s, int n) { return epp::CountFields(std::string(s, (size_t)n)); }
