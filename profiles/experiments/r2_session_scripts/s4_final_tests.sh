timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -n 6 | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -i "smoke\|error\|assert" | cut -c1-300
