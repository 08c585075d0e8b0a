#!/bin/bash
# Round-end validation + profile capture on one B200 (run through gpurun; outputs under gpurun_out/<tag>_*).
# Every ncu command runs only after the same command has exited 0 without ncu.
tag=${1:-z}
out=gpurun_out
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc $?"; tail -2 $out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc $?"
python bench.py --impl reference --steps 20 --warmup 3 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; echo "bench ref rc $?"
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc $?"
python tools/profile_step.py 1048576 6 > $out/${tag}_step_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tg_step_kernel -s 212 -c 2 -f -o $out/${tag}_step python tools/profile_step.py 1048576 6 > $out/${tag}_step_ncu.log 2>&1; echo "ncu step rc $?"
python tools/profile_render.py 4096 > $out/${tag}_render_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tg_render_stream -s 2 -c 1 -f -o $out/${tag}_render python tools/profile_render.py 4096 > $out/${tag}_render_ncu.log 2>&1; echo "ncu render rc $?"
python bench.py --steps 20 --warmup 5 --no-aux > $out/${tag}_bench_noaux.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 700 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 20 --warmup 5 --no-aux > $out/${tag}_launches.log 2>&1; echo "ncu launches rc $?"
python tools/bench_phases.py steady > $out/${tag}_phases.log 2>&1
python tools/bench_chunks.py 1048576 > $out/${tag}_chunks_1m.log 2>&1
python tools/bench_chunks.py 131072 > $out/${tag}_chunks_131k.log 2>&1
python tools/bench_render.py > $out/${tag}_render.log 2>&1
python tools/bench_sparse_host.py > $out/${tag}_sparse.log 2>&1
python tools/bench_sparse_host.py 131072 >> $out/${tag}_sparse.log 2>&1
nproc > $out/${tag}_nproc.txt
