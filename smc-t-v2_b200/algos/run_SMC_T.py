"""Mirror of the wiring part of src/algos/run_SMC_T.py (SMCTAlgo.__init__/_init_SMC_T, :15-110): builds the
SMC_Transformer from run.py-style arguments and installs the SMC parameters, so that `run.py -algo smc_t`
can be pointed at the B200 path.  Dataset objects, checkpoints, logging and the text-generation test
harness stay in the reference (out of scope, SURVEY.md §2 #9,#12)."""
from types import SimpleNamespace

from ..models.SMC_Transformer.SMC_Transformer import SMC_Transformer


def default_args(**kw):
    """run.py:35-94 defaults for the flags that feed the hot path."""
    a = dict(d_model=8, dff=8, num_layers=1, num_heads=1, pe=50, attn_w=None, full_model=True, p_drop=0., init_weights=1,
             smc=True, particles=10, sigmas=0.5, noise_dim="uni", EM_param=None, sigma_obs=0.5)
    a.update(kw)
    return SimpleNamespace(**a)


class SMCTAlgo:
    def __init__(self, dataset, args, output_size=None, seq_len=None):
        self.dataset = dataset
        self.output_size = output_size if output_size is not None else getattr(dataset, "output_size", 1)
        self.seq_len = seq_len if seq_len is not None else getattr(dataset, "seq_len", None)
        self.smc_transformer = SMC_Transformer(d_model=args.d_model, output_size=self.output_size, seq_len=self.seq_len,
                                               full_model=args.full_model, dff=args.dff,
                                               maximum_position_encoding=args.pe, attn_window=args.attn_w,
                                               num_layers=args.num_layers, num_heads=args.num_heads, rate=args.p_drop,
                                               init_weights=args.init_weights)
        self.distribution = args.smc
        self.particles = args.particles
        self.EM = args.EM_param is not None
        self.EM_param = args.EM_param
        self._init_SMC_T(args)

    def _init_SMC_T(self, args):
        """run_SMC_T.py:71-94."""
        if not args.smc:
            return
        if args.sigmas is None:
            raise ValueError("-sigmas is required with -smc True")
        if args.noise_dim == "multi":
            if not isinstance(args.sigmas, list):
                sigmas = [args.sigmas] * args.d_model
            elif len(args.sigmas) == args.d_model:
                sigmas = args.sigmas
            else:
                raise ValueError("Error in sigmas argument: should be either a scalar, either a list of length d_model args.")
        elif args.noise_dim == "noise_net":
            sigmas = args.noise_dim
        else:
            sigmas = args.sigmas
        dict_sigmas = dict(zip(["k", "q", "v", "z"], [sigmas for _ in range(4)]))
        self.smc_transformer.cell.add_SMC_parameters(dict_sigmas=dict_sigmas, num_particles=args.particles, EM=self.EM,
                                                     sigma_obs=getattr(args, "sigma_obs", None))
        assert self.smc_transformer.cell.noise == self.smc_transformer.cell.attention_smc.noise is True


algos = {"smc_t": SMCTAlgo}   # run.py:31
