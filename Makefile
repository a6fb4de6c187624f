# Convenience targets; the driver uses __graft_entry__.build() / smoke(), pytest and bench.py directly.
PY ?= python

build:            ## compile libphysad_b200.so (sm_100a), the CPU checkers and the reference test programs
	$(PY) -c "import __graft_entry__ as g; g.build()"

test-cpu: build   ## oracle pinned to the reference, boundary, partition, gloo 2-rank (no GPU needed)
	$(PY) -m pytest tests -q -m "not gpu"

test-gpu:         ## parity through the C-ABI on a B200
	$(PY) -m pytest tests -q -m gpu

bench:            ## headline benchmark, one JSON line
	$(PY) bench.py

golden:           ## regenerate tests/golden from the unmodified reference (needs /root/reference)
	$(PY) tests/golden/make_golden.py

clean:
	$(MAKE) -C phys_autodiff_b200/csrc clean
	$(MAKE) -C oracle clean
	$(MAKE) -C tests/refprogs clean

.PHONY: build test-cpu test-gpu bench golden clean
