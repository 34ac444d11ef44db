#!/bin/bash
# GPU validation of the training path
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_train_gpu.py -m gpu -x -q -s -k "training or loss_terms or train_gpu or trainer or v020_shaped_architecture_parameter or forward_in_training or second_order" 2>&1 | tail -40 | tee gpurun_out/pytest_train.log
