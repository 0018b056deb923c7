"""Drop-in for the subset of the reference's utils/tools.py that the model, test.py
and scripts/simple_train.py touch: `tools.abstract_config`,
`tools.abstract_model`, `tools.abs_test_model`, `tools.torch_warp`,
`tools.occ_check_model`, `tools.boundary_dilated_warp`, meters and timers.  Same names, arguments and
behaviour (citations per item); the warp runs on the library's kernel.

I/O, visualisation, augmentation and data loading (utils/tools.py:166-252,
:679-1207, :1341-1632) are out of scope."""
import time

import torch
import torch.nn as nn

from upflow_pytorch_b200 import ops


class tools():
    class abstract_config():
        """Attribute-bag configuration (utils/tools.py:32-107)."""
        name_filter_out_list = []

        def _public(self):
            skip = set(self.name_filter_out_list) | {'name_filter_out_list', 'get_name', 'update', 'update_ex_name',
                                                     'get_dict', 'check_length_of_file_path',
                                                     'check_length_of_file_name'}
            return sorted(n for n in dir(self) if '__' not in n and not n.startswith('_') and n not in skip)

        def get_name(self, print_now=True):
            names = self._public()
            if print_now:
                print('=' * 10)
                print('{')
                for n in names:
                    print("\t%-50s: '%s,', " % ("'%s'" % n, getattr(self, n)))
                print('}')
                print('=' * 10)
            return ''.join('%s|%s_' % (n, getattr(self, n)) for n in names)

        @classmethod
        def check_length_of_file_name(cls, file_name):
            return len(file_name) < 255

        @classmethod
        def check_length_of_file_path(cls, filepath):
            return len(filepath) < 4096

        def update(self, data: dict):
            # only attributes that already exist are set, each one reported (utils/tools.py:76-90)
            for n in dir(self):
                if not n.startswith('_') and n in data:
                    setattr(self, n, data[n])
                    print('set param ====  %s:   %s' % (n, data[n]))

        def get_dict(self):
            return {n: getattr(self, n) for n in dir(self) if not n.startswith('_')}

        def update_ex_name(self, ex_name: str):
            return ex_name

    class abstract_model(nn.Module):
        """state_dict save / (relaxed) load (utils/tools.py:109-155)."""

        def save_model(self, save_path):
            torch.save(self.state_dict(), save_path)

        def load_model(self, load_path, if_relax=False, if_print=True):
            if if_print:
                print('loading protrained model from %s' % load_path)
            loaded = torch.load(load_path, map_location=None if torch.cuda.is_available() else 'cpu')
            if if_relax:
                # keep only entries whose name AND shape match (utils/tools.py:115-125)
                own = self.state_dict()
                own.update({k: v for k, v in loaded.items() if k in own and v.shape == own[k].shape})
                self.load_state_dict(own)
            else:
                self.load_state_dict(loaded)

        @classmethod
        def choose_gpu(cls, model, gpu_opt=None):
            if gpu_opt is None:
                # the reference wraps the model in nn.DataParallel over every visible GPU (utils/tools.py:140): one
                # process, one thread per device.  This library's engine (CUDA graphs, per-shape workspaces) is built
                # for one process per GPU (upflow_pytorch_b200.train.Trainer / torchrun), so several devices in one
                # process are refused instead of run half-supported
                if torch.cuda.device_count() > 1:
                    raise RuntimeError("choose_gpu(gpu_opt=None) would wrap the model in nn.DataParallel over %d GPUs; "
                                       "upflow_pytorch_b200 runs one process per GPU: launch with torchrun and use "
                                       "upflow_pytorch_b200.train.Trainer, or pass gpu_opt=<index>" % torch.cuda.device_count())
                model = torch.nn.DataParallel(model.cuda(), device_ids=[0])
            elif gpu_opt == 0:
                model = model.cuda()
            else:
                if type(gpu_opt) != int:
                    raise ValueError('wrong gpu config, it show be int:  %s' % (str(gpu_opt)))
                torch.cuda.set_device(gpu_opt)
                model = model.cuda(gpu_opt)
            return model

        @classmethod
        def save_model_gpu(cls, model, path):
            if type(model).__name__ == torch.nn.DataParallel.__name__:
                model = model.module
            model.save_model(path)

    class abs_test_model():
        """Interface Evaluation_bench calls (utils/tools.py:157-164)."""
        save_dir = ''

        def eval_forward(self, im1, im2, gt, *args):
            return 0

        def eval_save_result(self, save_name, predflow, *args, **kwargs):
            pass

        def do_save_results(self, result_save_dir=None, some_save_results=False):
            self.save_dir = result_save_dir or ''

    class AverageMeter():
        def __init__(self):
            self.reset()

        def reset(self):
            self.val = self.avg = self.sum = self.count = 0

        def update(self, val, num):
            self.val = val
            self.sum += val * num
            self.count += num
            self.avg = self.sum / self.count

    class time_clock():
        def __init__(self):
            self.st = self.en = self.start_flag = 0

        def start(self):
            self.reset()
            self.start_flag = True
            self.st = time.time()

        def reset(self):
            self.start_flag = False
            self.st = self.en = 0

        def end(self):
            self.en = time.time()

        def get_during(self):
            return self.en - self.st

    @classmethod
    def tensor_gpu(cls, *args, check_on=True, gpu_opt=None, non_blocking=True):
        def move(a):
            if torch.is_tensor(a):
                return a.cuda(gpu_opt, non_blocking=non_blocking) if check_on else a.cpu()
            return a
        return [move(a) for a in args]

    @classmethod
    def check_tensor(cls, data, name, print_data=False, print_in_txt=None):
        if data.is_cuda:
            data = data.detach().cpu()
        a = data.numpy()
        print(name, 'shape', a.shape, 'max', a.max(), 'min', a.min(), 'mean', a.mean())

    @classmethod
    def torch_warp(cls, x, flo):
        """Warp x [B,C,H,W] by flo [B,2,H,W]; bilinear, zero padding, no validity mask (utils/tools.py:1274-1304)."""
        return ops.warp(x, flo, align_corners=False, use_mask=False)

    class boundary_dilated_warp():
        """Photometric-loss warp that samples the UN-CROPPED frame (utils/tools.py:350-499): the training crop starts
        at `start` inside the full image, so flows pointing outside the crop still find real pixels.  Loss-side op on
        3-channel images: one gather kernel each way for CUDA tensors (ops.boundary_warp);
        the index arithmetic and gathers in torch below are its A/B partner (`use_kernel = False`) and the CPU path of
        the oracle tests."""
        use_kernel = True

        @classmethod
        def get_grid(cls, batch_size, H, W, start):
            xx = torch.arange(0, W, device=start.device, dtype=torch.float32).view(1, 1, 1, W).expand(batch_size, 1, H, W)
            yy = torch.arange(0, H, device=start.device, dtype=torch.float32).view(1, 1, H, 1).expand(batch_size, 1, H, W)
            grid = torch.cat((xx, yy, torch.ones_like(xx)), 1)
            grid[:, :2] = grid[:, :2] + start
            return grid

        @classmethod
        def transformer(cls, I, vgrid, train=True):
            """Bilinear lookup of I [B,C,Hf,Wf] at absolute pixel positions vgrid [B,2,h,w]; corner indices are clamped
            to the image, weights are taken against the CLAMPED corners (utils/tools.py:383-470)."""
            B, C, Hf, Wf = I.shape
            x, y = vgrid[:, 0], vgrid[:, 1]                      # [B,h,w]
            x0 = torch.floor(x)
            y0 = torch.floor(y)
            x0c, x1c = x0.clamp(0, Wf - 1), (x0 + 1).clamp(0, Wf - 1)
            y0c, y1c = y0.clamp(0, Hf - 1), (y0 + 1).clamp(0, Hf - 1)
            flat = I.float().reshape(B, C, Hf * Wf)

            def take(yc, xc):
                idx = (yc.long() * Wf + xc.long()).reshape(B, 1, -1).expand(B, C, -1)
                return torch.gather(flat, 2, idx).reshape(B, C, *x.shape[1:])
            wa = ((x1c - x) * (y1c - y)).unsqueeze(1)
            wb = ((x1c - x) * (y - y0c)).unsqueeze(1)
            wc = ((x - x0c) * (y1c - y)).unsqueeze(1)
            wd = ((x - x0c) * (y - y0c)).unsqueeze(1)
            out = wa * take(y0c, x0c) + wb * take(y1c, x0c) + wc * take(y0c, x1c) + wd * take(y1c, x1c)
            return out if train else out.permute(0, 2, 3, 1)

        @classmethod
        def warp_im(cls, I_nchw, flow_nchw, start_n211):
            if cls.use_kernel and flow_nchw.is_cuda and not I_nchw.requires_grad and flow_nchw.shape[1] == 2:
                return ops.boundary_warp(I_nchw.float().to(flow_nchw.device), flow_nchw,
                                         start_n211.to(flow_nchw.device).float())               # csrc/loss.cu
            batch_size = I_nchw.shape[0]
            _, _, ph, pw = flow_nchw.shape
            grid = cls.get_grid(batch_size, ph, pw, start_n211.to(flow_nchw.device).float())
            return cls.transformer(I_nchw, grid[:, :2] + flow_nchw)

    class occ_check_model():
        """Forward/backward consistency occlusion masks (utils/tools.py:501-677); runs after the decoder on two
        2-channel flows, elementwise torch around the library warp."""

        def __init__(self, occ_type='for_back_check', occ_alpha_1=1.0, occ_alpha_2=0.05, sum_abs_or_squar=True,
                     obj_out_all='all'):
            self.occ_type_ls = ['for_back_check', 'forward_warp']
            assert occ_type in self.occ_type_ls
            assert obj_out_all in ['obj', 'out', 'all']
            self.occ_type = occ_type
            self.occ_alpha_1 = occ_alpha_1
            self.occ_alpha_2 = occ_alpha_2
            self.sum_abs_or_squar = True
            self.obj_out_all = obj_out_all

        def __call__(self, flow_f, flow_b, scale=1):
            if self.occ_type != 'for_back_check':
                raise ValueError('not implemented')
            if self.obj_out_all == 'out':
                return self.torch_outgoing_occ_check(flow_f), self.torch_outgoing_occ_check(flow_b)
            occ_1, occ_2 = self._forward_backward_occ_check(flow_f, flow_b, scale)
            if self.obj_out_all == 'all':
                return occ_1, occ_2
            return (self.torch_get_obj_occ_check(occ_1, self.torch_outgoing_occ_check(flow_f)),
                    self.torch_get_obj_occ_check(occ_2, self.torch_outgoing_occ_check(flow_b)))

        def _forward_backward_occ_check(self, flow_fw, flow_bw, scale=1):
            def mag(x):   # sum over channels of sqrt(x^2), i.e. |u|+|v| (utils/tools.py:556-560)
                return torch.sum(torch.pow(x ** 2, 0.5), dim=1, keepdim=True)
            mag_sq = mag(flow_fw) + mag(flow_bw)
            bw_warped = tools.torch_warp(flow_bw, flow_fw)
            fw_warped = tools.torch_warp(flow_fw, flow_bw)
            thresh = self.occ_alpha_1 * mag_sq + self.occ_alpha_2 / scale
            occ_fw = mag(flow_fw + bw_warped) < thresh
            occ_bw = mag(flow_bw + fw_warped) < thresh
            return occ_fw.float(), occ_bw.float()

        @classmethod
        def torch_outgoing_occ_check(cls, flow):
            B, C, H, W = flow.size()
            xx = torch.arange(0, W, device=flow.device, dtype=flow.dtype).view(1, 1, 1, W)
            yy = torch.arange(0, H, device=flow.device, dtype=flow.dtype).view(1, 1, H, 1)
            pos_x = xx + flow[:, 0:1]
            pos_y = yy + flow[:, 1:2]
            inside = (pos_x <= W - 1) & (pos_x >= 0) & (pos_y <= H - 1) & (pos_y >= 0)
            return inside.float()

        @classmethod
        def torch_get_obj_occ_check(cls, occ_mask, out_occ):
            return ((occ_mask == 1) | (out_occ == 0)).float()
