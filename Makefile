# Convenience targets; the driver uses __graft_entry__.build() / pytest / bench.py directly.
PY ?= python

build:            ## liblumacu.so (sm_100a) + oracle + facade test builds
	$(PY) -c "import __graft_entry__ as g; g.build()"

test-cpu:         ## everything that runs without a GPU
	$(PY) -m pytest tests -q -m "not gpu"

test-gpu:         ## parity tests proper (B200)
	$(PY) -m pytest tests -q -m gpu

bench:            ## headline benchmark, one GPU
	$(PY) bench.py

bench-reference:  ## the reference's own CPU implementation of the path
	$(PY) bench.py --impl reference

clean:
	$(MAKE) -C lumahdrv_b200/csrc clean
	$(MAKE) -C oracle clean
	$(MAKE) -C tests/cxx clean
	$(MAKE) -C tests/cxx -f Makefile.real clean

.PHONY: build test-cpu test-gpu bench bench-reference clean
