python tools/sanitize.py
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize.py > gpurun_out/san_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|sanitize pass" gpurun_out/san_$tool.log | sort | uniq -c | head -12
done
