// TEST INFRASTRUCTURE ONLY — configuration header handed to the *unmodified* reference
// sources under /root/reference through the reference's own out-of-source hook
// (config/config.hpp:13-24: `#ifdef MFEM_CONFIG_FILE / #include MFEM_CONFIG_FILE`).
// It replaces the `_config.hpp` the reference's build system would write; the values are
// the ones a serial CUDA build uses (config/config.hpp.in: MFEM_USE_CUDA), for the optional
// second baseline `make -C oracle refcuda`: the reference's OWN CUDA backend recompiled for sm_100. Nothing here is
// product code and nothing in the product may include it.
#ifndef B200PA_ORACLE_REF_CONFIG_CUDA_HPP
#define B200PA_ORACLE_REF_CONFIG_CUDA_HPP
#define MFEM_CONFIG_HEADER
#define MFEM_VERSION 40901
#define MFEM_VERSION_STRING "4.9.1"
#define MFEM_VERSION_TYPE ((MFEM_VERSION)%2)
#define MFEM_VERSION_TYPE_RELEASE 0
#define MFEM_VERSION_TYPE_DEVELOPMENT 1
#define MFEM_VERSION_MAJOR ((MFEM_VERSION)/10000)
#define MFEM_VERSION_MINOR (((MFEM_VERSION)/100)%100)
#define MFEM_VERSION_PATCH ((MFEM_VERSION)%100)
#define MFEM_SOURCE_DIR "/root/reference"
#define MFEM_INSTALL_DIR "/nonexistent"
#define MFEM_GIT_STRING "(unknown)"
#define MFEM_USE_DOUBLE
#define MFEM_USE_MEMALLOC
#define MFEM_USE_CUDA
#define MFEM_TIMER_TYPE 2
#endif
