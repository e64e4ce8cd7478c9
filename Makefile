# Convenience targets (everything is also reachable through the scripts they call).
PY ?= python

.PHONY: build test test-gpu test-cpp test-2gpu bench bench-ref docs census clean

build:            ## compile csrc/ for sm_100a into graphlearn_for_pytorch_b200/_ext/
	$(PY) -c "import __graft_entry__ as g; g.build()"

test: build       ## CPU tier (multi-process tests over localhost RPC / gloo)
	$(PY) -m pytest tests -q -m "not gpu"

test-gpu: build   ## needs one B200
	$(PY) -m pytest tests -q -m gpu

test-cpp:         ## native tests + ThreadSanitizer build of the shm ring
	scripts/run_cpp_ut.sh tsan

test-2gpu: build  ## needs two GPUs
	$(PY) -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/mp/p2p_check.py
	$(PY) tests/mp/feature_ipc_check.py

bench: build
	$(PY) bench.py --steps 50 --warmup 5

bench-ref:
	$(PY) bench.py --impl reference --steps 20 --warmup 3

docs:             ## regenerate the API reference
	$(PY) tools/gen_api_docs.py > docs/api.md

census:           ## SASS instruction census of the built extension
	$(PY) tools/sass_census.py > profiles/sass_summary.txt

clean:
	rm -rf graphlearn_for_pytorch_b200/_ext build *.egg-info
