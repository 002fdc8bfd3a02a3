# Common entry points (everything also works without make; see README.md)
PY ?= python
N  ?= 8

.PHONY: build test test-gpu bench bench-ref bench-all parity cov sass clean

build:            ## compile the sm_100a extension in place (no GPU needed)
	$(PY) -m torchdistpackage_b200.ops._build

test:             ## CPU / gloo suite
	$(PY) -m pytest tests -q -m "not gpu"

test-gpu:         ## on a B200 box
	$(PY) -m pytest tests -q -m gpu

bench:            ## flagship (BASELINE config #2) on one GPU
	$(PY) bench.py --gpus 1 --steps 10 --warmup 3

bench-ref:        ## the unmodified reference arm on one GPU
	$(PY) bench.py --impl reference --gpus 1 --steps 10 --warmup 3

bench-all:        ## N GPUs, headline + configs #3/#4/#5 for both arms
	$(PY) -m torch.distributed.run --nnodes=1 --nproc-per-node $(N) --master-addr 127.0.0.1 \
		bench.py --gpus $(N) --other-configs on

parity:           ## regenerate PARITY.md (file:line per row of the reference inventory)
	$(PY) scripts/gen_parity.py

cov:              ## which functions of the package the CPU suite never runs
	rm -rf build/cov && TDP_COV_DIR=build/cov $(PY) -m pytest tests -q -m "not gpu" && $(PY) tests/_cov.py build/cov

sass:             ## SASS listings per kernel family -> profiles/sass/
	bash scripts/sass_listings.sh

clean:
	rm -rf build torchdistpackage_b200/_C.so .pytest_cache .hypothesis
