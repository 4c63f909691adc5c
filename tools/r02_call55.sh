for d in 1 2 3; do for b in 1 2 8; do echo "== div $d B=$b"; SEUNET_BWD_WG_SMS_DIV=$d timeout 300 python tools/time_train.py $b 128 2>&1 | head -1; done; done
