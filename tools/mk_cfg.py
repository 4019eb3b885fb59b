import sys
sys.path.insert(0,'.')
import _pkg
pkg=_pkg.load()
open('gpurun_out/cfg2.cfg','w').write(pkg.cfg_text(3,['v1','v2'],2))
