#!/bin/bash
# Timing of experimental builds of the library (gym_treasure_game_b200/exp/libtg_<v>.so, built with extra -D flags and selected
# through TG_B200_LIB): usage exp_variants.sh <tag> <v> ...   Every run sits under `timeout`: an experimental kernel that hangs
# must not eat the GPU call.
out=gpurun_out
tag=${1:-ex}; shift
for v in "$@"; do
  lib=$PWD/gym_treasure_game_b200/exp/libtg_$v.so
  echo "== $v" >> $out/${tag}_steps.log
  TG_B200_LIB=$lib timeout 60 python tools/profile_step.py 1048576 12 >> $out/${tag}_steps.log 2>&1
  TG_B200_LIB=$lib timeout 60 python tools/profile_step.py 131072 12 >> $out/${tag}_steps.log 2>&1
done
cat $out/${tag}_steps.log
