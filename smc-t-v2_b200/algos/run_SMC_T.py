"""Mirror of src/algos/run_SMC_T.py (SMCTAlgo, :15-153) and of the parts of src/algos/generic.py (Algo, :12-66) it
relies on, so that `run.py -algo smc_t` can be pointed at the B200 path: `algos["smc_t"](dataset=dataset, args=args)`,
then `.train()` (run.py:128-131).  The model, the training step, the optimiser and the epoch loop run on the CUDA
path; the reference's control plane around them is reduced to what the path needs:
  * datasets: any object with the reference's protocol (`get_datasets()` + `data_to_dataset(train, val, test, num_dim)`
    returning iterables of (inputs, targets, attention_mask) batches, `output_size`, `seq_len`) or with ready-made
    `train_dataset` / `val_dataset` / `test_dataset` iterables;
  * logging: a std-lib logger writing `training_log.log` in the output folder when `args.output_path` is given;
  * checkpoints: `model.save_weights` / `load_weights` (npz) instead of tf.train.CheckpointManager;
  * `test()` (text generation, BLEU, GPT-2 perplexity: SURVEY.md §2 #9,#12) raises NotImplementedError."""
import datetime
import json
import logging
import os
from types import SimpleNamespace

import numpy as np

from ..models.SMC_Transformer.SMC_Transformer import SMC_Transformer
from ..models.SMC_Transformer.self_attention_SMC import logvar_vector
from ..train.optimizers import Adam, CustomSchedule
from ..train.train_functions import train_SMC_transformer


def default_args(**kw):
    """run.py:35-94 defaults for the flags that feed the hot path."""
    a = dict(d_model=8, dff=8, num_layers=1, num_heads=1, pe=50, attn_w=None, full_model=True, p_drop=0., init_weights=1,
             smc=True, particles=10, sigmas=0.5, noise_dim="uni", EM_param=None, sigma_obs=0.5,
             bs=32, ep=1, output_path=None, save_path=None, mc_samples=1, past_len=4, future_len=None,
             inference_resample=False, dataset="synthetic", algo="smc_t", temp=1., test_samples=None)
    a.update(kw)
    return SimpleNamespace(**a)


def _take1(dataset):
    for item in dataset:
        return item
    return None


class SMCTAlgo:
    def __init__(self, dataset, args, output_size=None, seq_len=None):
        # ---- Algo.__init__ (generic.py:12-27) ----
        self.dataset = dataset
        self.bs = getattr(args, "bs", None)
        self.EPOCHS = getattr(args, "ep", 0)
        self.output_path = getattr(args, "output_path", None)
        self.save_path = getattr(args, "save_path", None)
        self.start_epoch = 0
        self.mc_samples = getattr(args, "mc_samples", 1)
        self.past_len = getattr(args, "past_len", None)
        self.inference_resample = getattr(args, "inference_resample", False)
        # ---- SMCTAlgo.__init__ (run_SMC_T.py:16-47) ----
        self.lr = CustomSchedule(args.d_model)
        self.optimizer = Adam(self.lr, beta_1=0.9, beta_2=0.98, epsilon=1e-9)
        self.out_folder = self._create_out_folder(args)
        self.logger = self.create_logger()
        self.save_hparams(args)
        self.train_dataset, self.val_dataset, self.test_dataset = self.load_datasets(num_dim=4)
        if output_size is not None:
            self.output_size = output_size
        if seq_len is not None:
            self.seq_len = seq_len
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
        self.sigmas_after_training = None
        self.ckpt_manager = None
        if self.save_path is not None and os.path.exists(os.path.join(self.save_path, "model.npz")):
            self.smc_transformer.load_weights(os.path.join(self.save_path, "model"))
        if self.past_len is not None and self.seq_len is not None:
            assert self.past_len < self.seq_len, "past_len should be inferior to the sequence length of the dataset"
        fl = getattr(args, "future_len", None)
        self.future_len = fl if fl is not None else (None if self.past_len is None or self.seq_len is None else self.seq_len - self.past_len)

    # ---- folders / logging / hparams (reduced control plane) ----
    def _create_out_folder(self, args):
        """run_SMC_T.py:49-69; no folder is created when neither save_path nor output_path is given."""
        stamp = datetime.datetime.now().strftime("%Y%m%d-%H%M%S")
        if self.save_path is not None:
            out_folder = os.path.join(self.save_path, stamp)
        elif self.output_path is not None:
            out_file = "{}_{}_l{}_h{}_d{}".format(getattr(args, "dataset", "data"), getattr(args, "algo", "smc_t"),
                                                  args.num_layers, args.num_heads, args.d_model)
            if args.smc:
                out_file += "__p{}_sigmas_{}".format(args.particles, args.sigmas)
                if args.noise_dim == "multi":
                    out_file += "_multiD"
            out_folder = os.path.join(self.output_path, out_file, stamp)
        else:
            return None
        os.makedirs(out_folder, exist_ok=True)
        return out_folder

    def create_logger(self):
        logger = logging.getLogger("smct_b200.SMCTAlgo.%d" % id(self))
        logger.setLevel(logging.INFO)
        logger.propagate = False
        if self.out_folder is not None:
            logger.addHandler(logging.FileHandler(os.path.join(self.out_folder, "training_log.log")))
        else:
            logger.addHandler(logging.NullHandler())
        return logger

    def save_hparams(self, args):
        if self.out_folder is None:
            return
        with open(os.path.join(self.out_folder, "config.json"), "w") as fp:
            json.dump({k: str(v) for k, v in vars(args).items()}, fp, sort_keys=True, indent=4)

    def load_datasets(self, num_dim=4):
        """generic.py:50-66."""
        ds = self.dataset
        if ds is None:   # model wiring only (output_size / seq_len come from the constructor arguments)
            self.output_size, self.num_features, self.seq_len = 1, 1, None
            return None, None, None
        if hasattr(ds, "get_datasets"):
            train_data, val_data, test_data = ds.get_datasets()
            train, val, test = ds.data_to_dataset(train_data=train_data, val_data=val_data, test_data=test_data, num_dim=num_dim)
        else:
            train, val, test = ds.train_dataset, ds.val_dataset, getattr(ds, "test_dataset", None)
        self.output_size = getattr(ds, "output_size", 1)
        self.num_features = 1
        self.seq_len = getattr(ds, "seq_len", None)
        self.logger.info("output size: {}".format(self.output_size))
        self.logger.info("number of timeteps: {}".format(self.seq_len))
        return train, val, test

    # ---- SMC wiring ----
    def _init_SMC_T(self, args):
        """run_SMC_T.py:71-108 (incl. the smoke forward on the first training batch)."""
        if args.smc:
            self.logger.info("SMC Transformer for {} particles".format(args.particles))
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
            self.logger.info("Sigmas init: {}".format(dict_sigmas))
        first = _take1(self.train_dataset) if self.train_dataset is not None else None
        if first is not None:   # check the pass forward
            inp, tar = first[0], first[1]
            attn_mask = first[2] if len(first) > 2 else None
            (temp_preds, _), _, _ = self.smc_transformer(inputs=inp, targets=tar, attention_mask=attn_mask)
            self.logger.info("predictions shape: {}".format(tuple(temp_preds.shape)))

    # ---- run.py:128-135 ----
    def train(self):
        self.smc_transformer.training = True
        self.logger.info("hparams...")
        self.logger.info("num layers: {} - num_heads: {} - d_model: {} - batch size {} - full model? {} - dff: {} -attn window: {}".format(
            self.smc_transformer.num_layers, self.smc_transformer.num_heads, self.smc_transformer.d_model, self.bs,
            self.smc_transformer.full_model, self.smc_transformer.dff, self.smc_transformer.cell.attention_smc.attn_window))
        self.history = train_SMC_transformer(smc_transformer=self.smc_transformer, optimizer=self.optimizer, EPOCHS=self.EPOCHS,
                                             train_dataset=self.train_dataset, val_dataset=self.val_dataset,
                                             output_path=self.out_folder, ckpt_manager=self.ckpt_manager, logger=self.logger,
                                             start_epoch=self.start_epoch, num_train=1, EM_param=self.EM_param)
        att = self.smc_transformer.cell.attention_smc
        if self.distribution and att.noise_network is None:
            self.sigmas_after_training = {n: logvar_vector(getattr(att, "logvar_" + n)).detach().cpu().numpy() for n in ("k", "q", "v", "z")}
            if self.out_folder is not None:
                with open(os.path.join(self.out_folder, "logvar_after_training.json"), "w") as fp:
                    json.dump({k: str(np.squeeze(v)) for k, v in self.sigmas_after_training.items()}, fp)
        if self.out_folder is not None:
            self.smc_transformer.save_weights(os.path.join(self.out_folder, "model"))
        self.logger.info("-" * 60)
        return self.history

    def test(self, **kwargs):
        """run_SMC_T.py:319-470: text generation with GPT-2 perplexity / BLEU metrics."""
        raise NotImplementedError("SMCTAlgo.test (text generation, BLEU / GPT-2 perplexity metrics, run_SMC_T.py:319-470) is "
                                  "outside the particle-filter hot path (SURVEY.md §2 #9, #12); use "
                                  "smct_b200.eval.inference_functions for multi-step forecasting")


algos = {"smc_t": SMCTAlgo}   # run.py:31
