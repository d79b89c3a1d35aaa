# top-level build: the product (CUDA shim + C++ host library) and the checkers (oracle/)
all: product checkers

product:
	$(MAKE) -C biogpt.cpp_b200/csrc
	@if [ -f biogpt.cpp_b200/host/Makefile ]; then $(MAKE) -C biogpt.cpp_b200/host; fi

checkers:
	$(MAKE) -C oracle

clean:
	$(MAKE) -C biogpt.cpp_b200/csrc clean
	@if [ -f biogpt.cpp_b200/host/Makefile ]; then $(MAKE) -C biogpt.cpp_b200/host clean; fi
	$(MAKE) -C oracle clean

.PHONY: all product checkers clean
