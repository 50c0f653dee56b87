mkdir -p gpurun_out
( for mode in "api fence" "raw fence" "raw nofence"; do timeout 200 python tools/halves_timeline.py deep_sea/11 65536 $mode; done
  BSB_HOST_STAGE_ACTIONS=0 timeout 200 python tools/halves_timeline.py deep_sea/11 65536 raw nofence
  BSB_HOST_TIMING=1 timeout 200 python tools/e2e_timeline.py deep_sea/11 65536 ) > gpurun_out/halves_timeline.txt 2>&1
cat gpurun_out/halves_timeline.txt
