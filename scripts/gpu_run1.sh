set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30
timeout 300 python - <<'PY'
import json
from vtm_b200.peaks import all_peaks
for r in all_peaks(8192): print(json.dumps(r))
PY
