for v in onefwd twofwd; do echo "== $v"; B200CA_LIB=/root/repo/pyca_b200/variants/$v.so timeout 200 python tools/batch_probe.py 2>&1 | grep -E "B=24|B=96"; done
