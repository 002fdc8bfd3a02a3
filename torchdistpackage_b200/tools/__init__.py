from .module_profiler import register_profile_hooks, report_prof, get_model_profile, remove_profile_hooks
from .module_replace import replace_all_module
from .debug_nan import (check_tensors, check_model_params, fwd_hook_wrapper, bwd_hook_wrapper,
                        register_nan_hooks)
from .watchdog import StepWatchdog, HANG_EXIT_CODE
from .metrics import MetricsLogger
from .int8_linear import Int8WeightOnlyLinear, replace_linear_by_int8
from .bnb_fc import replace_linear_by_bnb
from .bminf_int8 import replace_linear_by_bminf
