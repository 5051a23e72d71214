for v in base ldg noperm; do echo "== $v"; B200CA_LIB=/root/repo/pyca_b200/variants/$v.so timeout 200 python tools/symm_probe.py 2>&1 | grep "p2p-halo"; done
