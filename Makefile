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
