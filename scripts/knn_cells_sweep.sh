for c in 2 4 8 16; do echo "== cells/pt $c"; AOPT_KNN_CELLS_PER_POINT=$c python scripts/knn_scale_sweep.py 1.0 1.25 1.6 2>&1 | tail -8; done
