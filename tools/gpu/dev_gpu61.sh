#!/bin/bash
# artefact writers + preproc_path ingest (row f-3), distortion report (row f-4), drop-ins
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_artefacts.py tests/test_metrics.py tests/test_dropin_gpu.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/artefacts_tests.log
cat gpurun_out/artefacts_tests.log
