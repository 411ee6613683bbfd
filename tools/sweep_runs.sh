for rs in 3 2 1; do for gss in 0 2 8; do echo "rs=$rs gss=$gss"; RT_RUN_SHIFT=$rs RT_GSS=$gss RT_STRIPS=1 python tools/quick_bench.py 40 2>&1 | tail -1; RT_RUN_SHIFT=$rs RT_GSS=$gss python tools/times_shard.py 8 2>&1 | tail -2 | head -1; done; done
echo auto; RT_STRIPS=1 python tools/quick_bench.py 40 2>&1 | tail -1; python tools/times_shard.py 8 2>&1 | tail -2
