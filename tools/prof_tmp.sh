python -m pytest tests -m gpu -x -q -k "camera or golden or example" 2>&1 | tail -3
python tools/render_probe.py from_the_readme 4096 2>&1 | grep -v Warn; python tools/render_probe.py basic_env 4096 2>&1 | grep -v Warn
