# oracle/ref.mk -- TEST INFRASTRUCTURE ONLY.
#
# Compiles the UNMODIFIED reference (AthenaK, /root/reference) into oracle/_ref/ straight from
# the sources where they lie, with g++ and this short Makefile (the reference's own CMake build
# is NOT run).  The only third-party piece is the vendored Kokkos 4.7.2, pulled in through
# Kokkos' own embedded GNU-make fragment (kokkos/Makefile.kokkos), which also writes
# KokkosCore_config.h into the build directory.  Nothing is copied into the repo; all outputs
# go to oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun).
#
# Flags follow the reference Release build (-O3 -DNDEBUG -std=c++17, no -march, hence no FMA
# contraction: SURVEY.md section 0 fact 3) plus -fopenmp: the Kokkos::OpenMP backend is
# bitwise identical to Kokkos::Serial on this path, so one binary
# serves both as parity oracle (OMP_NUM_THREADS=1) and as the all-cores CPU baseline.
#
# usage (see oracle/build_ref.sh):
#   make -C oracle/_ref/build_<name> -f <repo>/oracle/ref.mk PROBLEM=<pgen> NAME=<name> -j8

REF      ?= /root/reference
PROBLEM  ?= built_in_pgens
NAME     ?= athena_builtin
OUTDIR   ?= $(abspath ..)

SRCDIR   := $(REF)/src
# KDEV/KARCH/RCXX select the Kokkos backend: the default is the CPU oracle; the GPU baseline
# ("the reference's own Kokkos-CUDA kernels on the same B200") uses
#   KDEV=Cuda,Serial KARCH=Blackwell100 RCXX=$(REF)/kokkos/bin/nvcc_wrapper ROMP=
KDEV     ?= OpenMP,Serial
KARCH    ?=
RCXX     ?= g++
ROMP     ?= -fopenmp
CXX      := $(RCXX)
CXXFLAGS := -O3 -DNDEBUG -std=c++17 $(ROMP)

KOKKOS_PATH    := $(REF)/kokkos
KOKKOS_DEVICES := $(KDEV)
KOKKOS_ARCH    := $(KARCH)
KOKKOS_OPTIONS := aggressive_vectorization
KOKKOS_USE_DEPRECATED_MAKEFILES := 1
KOKKOS_CXX_STANDARD := c++17

default: $(OUTDIR)/$(NAME)

include $(KOKKOS_PATH)/Makefile.kokkos

# source list = the add_executable() list of the reference (src/CMakeLists.txt:8-198)
ATHENA_SRC := $(shell sed -n '/add_executable(/,/^)/p' $(SRCDIR)/CMakeLists.txt | grep -o '[A-Za-z0-9_/.-]*\.cpp')
ifneq ($(PROBLEM),built_in_pgens)
  ATHENA_SRC += pgen/$(PROBLEM).cpp
  USER_PROBLEM_ENABLED := 1
else
  USER_PROBLEM_ENABLED := 0
endif
ATHENA_OBJ := $(addprefix obj/,$(ATHENA_SRC:.cpp=.o))

# config.hpp is normally produced by CMake from config.hpp.in (CMakeLists.txt:34-131)
config.hpp: $(REF)/config.hpp.in
	sed -e 's|@PROBLEM@|$(PROBLEM)|' -e 's|@USER_PROBLEM_ENABLED@|$(USER_PROBLEM_ENABLED)|' \
	    -e 's|@TWO_PUNCTURES@|0|' -e 's|@athena_VERSION_MAJOR@|1|' -e 's|@athena_VERSION_MINOR@|0|' \
	    -e 's|@SINGLE_PRECISION_ENABLED@|0|' -e 's|@MPI_PARALLEL_ENABLED@|0|' \
	    -e 's|@OPENMP_PARALLEL_ENABLED@|$(if $(ROMP),1,0)|' -e 's|@COMPILER_CHOICE@|g++|' \
	    -e 's|@COMPILER_COMMAND@|g++|' -e 's|@COMPILER_FLAGS@|$(CXXFLAGS)|' \
	    -e 's|@PAR_LOOP_LAYOUT@|MDRANGE_LOOP|' -e 's|@PAR_LOOP_INNER_LAYOUT@|INNER_TTR_LOOP|' $< > $@

obj/%.o: $(SRCDIR)/%.cpp config.hpp $(KOKKOS_CPP_DEPENDS)
	@mkdir -p $(dir $@)
	$(CXX) $(KOKKOS_CPPFLAGS) $(KOKKOS_CXXFLAGS) $(CXXFLAGS) -I. -I$(SRCDIR) -c $< -o $@

$(OUTDIR)/$(NAME): $(ATHENA_OBJ) $(KOKKOS_LINK_DEPENDS)
	$(CXX) $(KOKKOS_LDFLAGS) $(CXXFLAGS) $(ATHENA_OBJ) $(KOKKOS_LIBS) -lpthread -o $@
