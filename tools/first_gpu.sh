set -x
nvidia-smi -L; nproc; ls /usr/lib/x86_64-linux-gnu/libnvoptix* 2>&1 | head -2; find / -name 'optix.h' -not -path '*/proc/*' 2>/dev/null | head -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30
