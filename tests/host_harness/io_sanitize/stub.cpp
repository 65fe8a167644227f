// minimal stand-ins for the two library internals pct_io.cu uses, so that it can be built alone with sanitizers
#include <string>
namespace pct { static thread_local std::string g_err; void set_error(const std::string& m) { g_err = m; } }
extern "C" const char* pct_last_error(void) { return pct::g_err.c_str(); }
