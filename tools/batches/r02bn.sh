cd $GRAFT_REPO_ROOT
P=tools/sustained_probe.sh
for n in 10 30 100 300; do $P "bwd x$n" build/fa_check time 512 4096 128 1 $n 1; sleep 3; done
for n in 10 100 1000; do $P "fwd x$n" build/fa_check time 512 4096 128 1 $n 0; sleep 3; done
$P "bwd non-causal x150" build/fa_check time 512 4096 128 0 150 1; sleep 3
FA_B200_BWD_DEBUG_DQ=sk LD_LIBRARY_PATH=build/variants/dbg $P "bwd no reduce-adds x300" build/fa_check time 512 4096 128 1 300 1; sleep 3
FA_B200_BWD_DEBUG_DQ=st LD_LIBRARY_PATH=build/variants/dbg $P "bwd plain stores x300" build/fa_check time 512 4096 128 1 300 1; sleep 3
LD_LIBRARY_PATH=build/variants/dbg $P "bwd dbg-lib x300" build/fa_check time 512 4096 128 1 300 1; sleep 3
$P "bwd dropout x60" build/fa_check time 128 16384 128 1 60 1 0.1; sleep 3
$P "fwd d64 x1000" build/fa_check time 512 4096 64 1 1000 0
