// Stub of the build-generated config header (reference: src/config/4C_config.hpp.in:12-33).
// Only the macros the header-only numerics need. Test infrastructure, not product code.
#ifndef FOUR_C_CONFIG_HPP
#define FOUR_C_CONFIG_HPP
#define FOUR_C_VERSION_FULL "oracle-ref-shim"
#define FOUR_C_NAMESPACE_OPEN namespace FourC {
#define FOUR_C_NAMESPACE_CLOSE }
#define FOUR_C_HIDDEN __attribute__((visibility("hidden")))
#endif
