# Builds the OpenABL compiler driver (front end + cuda backend).
CXX ?= g++
CXXFLAGS ?= -O2 -std=c++14 -Wall -Wextra -Wno-unused-parameter
SRCS := src/main.cpp src/Parser.cpp src/Sema.cpp src/FileUtil.cpp $(wildcard src/backend/*.cpp)
HDRS := $(wildcard src/*.hpp src/backend/*.hpp)

OpenABL: $(SRCS) $(HDRS)
	$(CXX) $(CXXFLAGS) -o $@ $(SRCS)

clean:
	rm -f OpenABL
.PHONY: clean

# ---- device runtime (asset/cuda/libabl_cuda.so), cross-compiled for sm_100a ---------------
NVCC ?= nvcc
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --cudart shared -Xcompiler -fPIC -Iinclude
RT_SRCS := asset/cuda/abl_runtime.cu

runtime: asset/cuda/libabl_cuda.so
asset/cuda/libabl_cuda.so: $(RT_SRCS) include/abl_cuda.h
	$(NVCC) $(NVFLAGS) -shared -o $@ $(RT_SRCS) -lnccl

all: OpenABL runtime
.PHONY: runtime all
