echo "== default"; timeout 120 python tools/r02d_dbg.py cfg1_concat_eval 2>&1 | grep -v Warn | tail -12
echo "== PDL off"; FB200_PDL=0 timeout 120 python tools/r02d_dbg.py cfg1_concat_eval 2>&1 | grep -v Warn | tail -12
echo "== one stream"; timeout 120 python tools/r02d_dbg.py cfg1_concat_eval $(python -c "
import sys; sys.path.insert(0,'multimodal-model-skin-lesion-classifier_b200')
from fusion_b200 import _lib; print(_lib.FLAG_ONE_STREAM)") 2>&1 | grep -v Warn | tail -12
echo "== csplit cap 2"; FB200_TC_CSPLIT=2 timeout 120 python tools/r02d_dbg.py cfg1_concat_eval 2>&1 | grep -v Warn | tail -12
