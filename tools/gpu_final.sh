#!/bin/bash
bash tools/gpu_round.sh
bash tools/gpu_prof_all.sh
