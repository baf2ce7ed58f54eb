"""``photon.utils`` is one module in the reference; here it is a package (``core`` = that module's role, ``flat`` = flat
parameter storage, ``hw`` = device facts, ``trace`` = Chrome-trace / NVTX spans). The names of ``core`` are re-exported so
``from photon_b200.utils import get_parameters_from_state`` works like ``from photon.utils import …`` did."""
from photon_b200.utils.core import *  # noqa: F401,F403
from photon_b200.messages import ClientState  # noqa: F401 - lives with the other wire types here; ``photon.utils.ClientState`` in the reference
from photon_b200.utils.core import (clean_parameter_name, construct_parameters_dict,  # noqa: F401
                                    dump_model_parameters_to_file, get_list_of_parameters_names, get_parameters_from_state,
                                    get_trainable_params_dict, l2_norm, load_model_parameters_from_file, parameters_checker,
                                    set_trainer_params_from_ndarrays, set_trainer_trainable_params_dict, sum_of_squares, wandb_init)
from photon_b200.utils.core import (NoOpContextManager, custom_ray_garbage_collector, download_file_from_s3,  # noqa: F401
                                    get_unigram_probabilities_tensor, merge_freq_dicts, set_parameters, upload_file_to_s3)
