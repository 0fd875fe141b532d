timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d[\"ms_per_step\"], d[\"roofline\"][\"phase_ms_median\"][\"lik_ms\"], d[\"elbo\"])"
