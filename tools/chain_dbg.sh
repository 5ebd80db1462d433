cd tests
for c in conv_chain_1tile conv_chain_ragged_repeat conv_chain_few_ctas conv_chain_24_layers conv_chain_two_tiles_per_cta conv_chain_full; do
  timeout 60 python -c "
import gpu_checks as g
r = g.CHECKS['$c']()
print('$c', 'OK')
" 2>&1 | grep -v "^$" | grep -i "OK\|timeout\|error\|tg_conv" | head -5
done
timeout 200 compute-sanitizer --tool memcheck --print-limit 5 python -c "
import gpu_checks as g
g.CHECKS['conv_chain_two_tiles_per_cta']()
" 2>&1 | grep -v "^$" | head -40
