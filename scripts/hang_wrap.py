import faulthandler, sys, runpy, os
faulthandler.dump_traceback_later(50, exit=True)
sys.argv = sys.argv[1:]
runpy.run_path(sys.argv[0], run_name="__main__")
