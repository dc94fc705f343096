# Builds libdigat_sm100.so (sm_100a only) in-tree.  `python -c "import __graft_entry__ as g; g.build()"` calls this recipe.
NVCC ?= nvcc
NVCCFLAGS ?= -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC
SRC := digat_b200/csrc/digat_abi.cu
HDR := $(wildcard digat_b200/csrc/*.cuh) include/digat_sm100.h
OUT := digat_b200/libdigat_sm100.so

$(OUT): $(SRC) $(HDR)
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(SRC)

ptxas-info: $(SRC) $(HDR)
	$(NVCC) $(NVCCFLAGS) -Xptxas -v -shared -o /tmp/digat_ptxas_info.so $(SRC)

clean:
	rm -f $(OUT)
.PHONY: clean ptxas-info
