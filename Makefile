# Builds libyolov6_b200.so (sm_100a only) and the C part of the oracle.  `python -c "import __graft_entry__ as g; g.build()"` calls this.
NVCC ?= nvcc
NVFLAGS = -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
SRC = $(wildcard yolov6_b200/csrc/*.cu)
HDR = $(wildcard yolov6_b200/csrc/*.cuh yolov6_b200/csrc/*.h include/*.h)
LIB = yolov6_b200/libyolov6_b200.so

all: $(LIB)

build/%.o: yolov6_b200/csrc/%.cu $(HDR)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(patsubst yolov6_b200/csrc/%.cu,build/%.o,$(SRC))
	$(NVCC) $(NVFLAGS) -shared -o $@ $^

clean:
	rm -rf build $(LIB)
.PHONY: all clean
