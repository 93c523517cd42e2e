run() { tool=$1; shift; echo "== $*"; timeout 240 compute-sanitizer --tool $tool python tools/sanitize_run.py "$@" 2>&1 | grep -v "^=========\s*$" | tail -4; }
for tool in memcheck racecheck; do
  {
  run $tool dpdfnet2
  run $tool dpdfnet2 overlap=2
  run $tool dpdfnet2 intra_frag=0 intra_sr=1 intra_dup=4
  run $tool dpdfnet2 intra_frag=0 intra_sr=1 intra_dup=2
  run $tool dpdfnet2_48khz_hr
  run $tool dpdfnet2 dfp_early=1
  } > gpurun_out/r4F_sanitizer_$tool.log 2>&1
done
tail -3 gpurun_out/r4F_sanitizer_memcheck.log; tail -3 gpurun_out/r4F_sanitizer_racecheck.log
