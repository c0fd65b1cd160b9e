"""Epoch-indexed hyper-parameter schedule for the preconditioner, with the
semantics of the reference's ``KFACParamScheduler``
(kfac/kfac_preconditioner_base.py:233-301): multiplicative factors applied to
``param_groups[0]['damping' | 'fac_update_freq' | 'kfac_update_freq']`` at the
listed epochs, ``int()`` truncation of the frequencies, picked up by the next
``step()``."""
from __future__ import annotations


class KFACParamScheduler:
    def __init__(self, kfac, damping_alpha=1, damping_schedule=None, update_freq_alpha=1,
                 update_freq_schedule=None, start_epoch=0):
        self.kfac = kfac
        params = self.kfac.param_groups[0]
        self.damping_base = params['damping']
        self.damping_alpha = damping_alpha
        self.damping_schedule = damping_schedule
        self.damping_factor_func = self._get_factor_func(damping_schedule, damping_alpha)
        self.fac_update_freq_base = params['fac_update_freq']
        self.kfac_update_freq_base = params['kfac_update_freq']
        self.update_freq_alpha = update_freq_alpha
        self.update_freq_schedule = update_freq_schedule
        self.update_freq_factor_func = self._get_factor_func(update_freq_schedule, update_freq_alpha)
        self.epoch = start_epoch

    @staticmethod
    def _get_factor_func(schedule, alpha):
        milestones = sorted(schedule, reverse=True) if schedule is not None else []
        if schedule is not None:
            schedule.sort(reverse=True)        # the reference sorts the caller's list in place

        def factor_func(epoch):
            factor = 1.
            for e in milestones:
                if epoch >= e:
                    factor *= alpha
            return factor
        return factor_func

    def step(self, epoch=None):
        if epoch is not None:
            self.epoch = epoch
        else:
            self.epoch += 1
        params = self.kfac.param_groups[0]
        params['damping'] = self.damping_base * self.damping_factor_func(self.epoch)
        factor = self.update_freq_factor_func(self.epoch)
        params['fac_update_freq'] = int(self.fac_update_freq_base * factor)
        params['kfac_update_freq'] = int(self.kfac_update_freq_base * factor)
