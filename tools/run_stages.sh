for s in 4 5 6 8; do echo "== items per CTA $s"; G2P_WIN_ITEMS_PER_CTA=$s G2P_ONLY=marked python tools/mask_variants.py 2>&1 | tail -2; done
